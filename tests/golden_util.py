import os

import oracle_lib as ol
import synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")


def golden_cases():
    return sorted(d for d in os.listdir(GOLD) if os.path.isdir(os.path.join(GOLD, d)))


def load_case(name):
    d = os.path.join(GOLD, name)
    c = dict(dir=d, iu=os.path.join(d, "index_u.bin1"), id=os.path.join(d, "index_d.bin2"),
             map=os.path.join(d, "genome_map.out"), fq=os.path.join(d, "reads.fq"))
    c["reads"] = synth.read_fastq(c["fq"])
    c["dump"] = {m: synth.parse_ref_dump(os.path.join(d, "dump_%s.txt" % m)) for m in ("p", "sc")}
    c["G"] = c["dump"]["p"]["g"]
    c["bases"], c["offsets"], c["lengths"] = ol.pack_reads(c["reads"])
    return c
