"""Parallel FASTQ ingest of the CLI (cammiq_b200/csrc/host/fastq_reader.cpp)
against the reference's reader, FqReader::readFastq (query.cpp:371-425):

  * the REAL reader: `oracle/_ref/ref_harness readdump` runs the unmodified readFastq with srand()
    interposed, so that CAMMIQ_SEED seeds the N substitution on both sides (tests marked `ref`);
  * a line-by-line Python restatement of it (records by line number, (uint8_t) length,
    min-length filter, one rand() per accepted read), itself checked against the real reader.

Runs without a GPU (`cammiq --dump_reads`)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "cammiq_b200", "cammiq")
REF_HARNESS = os.path.join(REPO, "oracle", "_ref", "ref_harness")


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "cammiq_b200", "csrc"), "../cammiq"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return CLI


def reference_reader(text, min_len, seed):
    """What the reference holds after readFastq, as (length, bases or None when invalid)."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(seed)
    lines = text.split(b"\n")
    final_newline = bool(lines) and lines[-1] == b""
    if final_newline:
        lines.pop()                      # getline does not produce a line after the final newline
    out = []
    i = 0
    while i < len(lines):                # header
        # a header that is the last line: with a final newline the second getline extracts nothing
        # (empty bases); without one the stream is already at eof and `bases` keeps the header text
        bases = lines[i + 1] if i + 1 < len(lines) else (b"" if final_newline else lines[i])
        if len(bases) >= min_len:
            sub = b"ACGT"[libc.rand() & 3]
            r = bytes(sub if ch == ord("N") else ch for ch in bases)
            ln = len(r) & 0xFF
            r = r[:ln]
            ok = all(ch in b"ACGTacgt" for ch in r)
            out.append((ln, r.upper() if ok else None))
        i += 4
    return out


def run_reader(cli, path, min_len, seed, threads):
    env = dict(os.environ, CAMMIQ_SEED=str(seed), CAMMIQ_IO_THREADS=str(threads))
    p = subprocess.run([cli, "--dump_reads", path, str(min_len)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, check=True)
    got = []
    for line in p.stdout.split(b"\n")[:-1]:
        ln, _, bases = line.partition(b" ")
        got.append((int(ln), bases))
    return got, p.stderr.decode()


def make_fastq(rng, n, with_n=True, crlf_every=0, ragged=True, trailing_newline=True):
    recs = []
    for i in range(n):
        ln = int(rng.integers(0, 300)) if ragged else 100
        seq = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), ln).tobytes())
        if with_n and ln and i % 5 == 0:
            for pos in rng.integers(0, ln, 3):
                seq[int(pos)] = ord("N")
        if ln and i % 37 == 0:
            seq[int(rng.integers(0, ln))] = ord("n")      # lowercase n is NOT substituted by the reference
        if ln and i % 41 == 0:
            seq = bytearray(bytes(seq).lower())
        qual = bytes(rng.choice(np.frombuffer(b"@+IJ#5", dtype=np.uint8), ln).tobytes())  # '@' and '+' inside qualities
        if crlf_every and i % crlf_every == 0:
            seq += b"\r"
        recs.append(b"@read%d\n%s\n+\n%s" % (i, bytes(seq), qual))
    text = b"\n".join(recs)
    return text + (b"\n" if trailing_newline else b"")


def real_reference_reader(path, min_len, seed):
    """The unmodified FqReader::readFastq through ref_harness (srand interposed: CAMMIQ_SEED)."""
    env = dict(os.environ, CAMMIQ_SEED=str(seed))
    p = subprocess.run([REF_HARNESS, "readdump", path, str(min_len)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, check=True)
    out = []
    for line in p.stdout.split(b"\n")[:-1]:
        ln, _, bases = line.partition(b" ")
        ok = all(ch in b"ACGTacgt" for ch in bases)
        out.append((int(ln), bases.upper() if ok else None))
    return out


def compare(cli, tmp_path, text, min_len, seed, threads):
    path = str(tmp_path / "x.fq")
    with open(path, "wb") as f:
        f.write(text)
    want = reference_reader(text, min_len, seed)
    if os.access(REF_HARNESS, os.X_OK):
        # pin the restatement (and through it the CLI) to the real reader (CR bytes included: the
        # reference keeps the CR of a CRLF line as the read's last byte, which makes the read invalid)
        assert real_reference_reader(path, min_len, seed) == want
    got, err = run_reader(cli, path, min_len, seed, threads)
    assert len(got) == len(want), err
    for i, ((gl, gb), (wl, wb)) in enumerate(zip(got, want)):
        if wb is None:
            assert gl == 0, (i, gl)                      # invalid read: handed to the kernel with length 0
        else:
            assert (gl, gb) == (wl, wb), i
    return err


@pytest.mark.parametrize("threads", [1, 2, 7])
def test_ragged_reads_with_n_and_min_length(cli, tmp_path, threads):
    rng = np.random.default_rng(threads)
    text = make_fastq(rng, 40000)                         # ~12 MB: several slices per thread count
    for min_len in (0, 50):
        err = compare(cli, tmp_path, text, min_len, 11, threads)
    assert "%d threads" % threads in err


def test_fixed_length_no_trailing_newline_and_crlf(cli, tmp_path):
    rng = np.random.default_rng(9)
    compare(cli, tmp_path, make_fastq(rng, 30000, ragged=False, trailing_newline=False), 0, 3, 4)
    compare(cli, tmp_path, make_fastq(rng, 30000, ragged=False, crlf_every=100), 0, 3, 4)


def test_truncated_and_tiny_files(cli, tmp_path):
    compare(cli, tmp_path, b"", 0, 1, 4)
    compare(cli, tmp_path, b"@only_header\n", 0, 1, 4)            # bases line missing: an empty read
    compare(cli, tmp_path, b"@only_header", 0, 1, 4)
    compare(cli, tmp_path, b"@r\nACGTN\n", 0, 1, 4)               # record cut after the bases
    compare(cli, tmp_path, b"@r\nACGTN\n+\nIIIII\n@s\nAC", 0, 1, 4)
    compare(cli, tmp_path, b"@r\nACGTN\n+\nIIIII\n@s\nAC", 3, 1, 4)


@pytest.mark.ref
@pytest.mark.parametrize("min_len", [0, 60])
def test_against_the_real_reference_reader(cli, tmp_path, min_len):
    """Reads with N (substituted by the reference's own rand() sequence), 300-bp reads (length
    wraps to uint8), lower case, truncated last records -- the CLI's reads must equal what the
    unmodified readFastq holds."""
    rng = np.random.default_rng(77 + min_len)
    for k, text in enumerate((make_fastq(rng, 20000), make_fastq(rng, 5000, trailing_newline=False),
                              make_fastq(rng, 3000) + b"@cut_after_header", make_fastq(rng, 3000) + b"@cut_after_header\n",
                              make_fastq(rng, 3000) + b"@r\nACGTNNACGT")):
        path = str(tmp_path / ("r%d.fq" % k))
        with open(path, "wb") as f:
            f.write(text)
        want = real_reference_reader(path, min_len, 5 + k)
        got, err = run_reader(cli, path, min_len, 5 + k, 4)
        assert len(got) == len(want), (k, err)
        n_subst = 0
        for i, ((gl, gb), (wl, wb)) in enumerate(zip(got, want)):
            if wb is None:
                assert gl == 0, (k, i)
            else:
                assert (gl, gb) == (wl, wb), (k, i)
                n_subst += 1
        assert n_subst > 0
