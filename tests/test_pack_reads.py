"""Host-side 2-bit read packer (cq_pack_reads) against a numpy restatement of
the layout include/cammiq_gpu.h documents.  No GPU needed: the packer is host code."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cammiq_b200 as cq

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = np.full(256, 255, dtype=np.uint8)
for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
    CODE[ch] = v


def numpy_pack(read):
    """bytes -> (packed bytes, valid) per the documented layout: base j -> byte j/4, bits 7-2(j%4).."""
    codes = CODE[np.frombuffer(read, dtype=np.uint8)]
    valid = not (codes == 255).any()
    pad = (-len(codes)) % 4
    c = np.concatenate([codes & 3, np.zeros(pad, np.uint8)]).reshape(-1, 4).astype(np.uint32)
    return ((c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]).astype(np.uint8), valid


def make_reads(rng, n, max_len=255, bad_every=0):
    reads = []
    for i in range(n):
        ln = int(rng.integers(0, max_len + 1))
        r = bytearray(rng.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), ln).tobytes())
        if bad_every and i % bad_every == 0 and ln:
            r[int(rng.integers(0, ln))] = rng.choice(np.frombuffer(b"NnXR-*@\x00\xff", dtype=np.uint8))
        reads.append(bytes(r))
    return reads


def check(reads, threads, offsets_mode):
    lengths = np.array([len(r) for r in reads], dtype=np.uint8)
    if offsets_mode:
        offs = np.zeros(len(reads), dtype=np.uint64)
        offs[1:] = np.cumsum([len(r) + 3 for r in reads])[:-1]       # gaps between reads
        bases = np.full(int(offs[-1]) + len(reads[-1]) + 3, ord("N"), dtype=np.uint8)
        for o, r in zip(offs, reads):
            bases[int(o):int(o) + len(r)] = np.frombuffer(r, dtype=np.uint8)
        packed, out_len, bad = cq.pack_reads(bases, offs, lengths, threads=threads)
    else:
        stride = 256
        bases = np.full(len(reads) * stride, ord("N"), dtype=np.uint8)
        for i, r in enumerate(reads):
            bases[i * stride:i * stride + len(r)] = np.frombuffer(r, dtype=np.uint8)
        packed, out_len, bad = cq.pack_reads(bases, None, lengths, stride=stride, threads=threads)
    n_bad = 0
    for i, r in enumerate(reads):
        want, valid = numpy_pack(r)
        if valid:
            assert out_len[i] == len(r)
            assert np.array_equal(packed[i, :len(want)], want), (i, len(r))
        else:
            assert out_len[i] == 0
            n_bad += 1
    assert bad == n_bad


@pytest.mark.parametrize("threads", [1, 3, 8])
@pytest.mark.parametrize("offsets_mode", [False, True])
def test_pack_matches_numpy(threads, offsets_mode):
    rng = np.random.default_rng(7 + threads)
    check(make_reads(rng, 3000, bad_every=17), threads, offsets_mode)


def test_every_length_and_tail_padding_is_zero():
    reads = [b"T" * n for n in range(0, 256)]            # all-ones codes: padding bits must stay 0
    lengths = np.array([len(r) for r in reads], dtype=np.uint8)
    bases = np.zeros(256 * 256, dtype=np.uint8)
    for i, r in enumerate(reads):
        bases[i * 256:i * 256 + len(r)] = np.frombuffer(r, dtype=np.uint8)
        bases[i * 256 + len(r):(i + 1) * 256] = ord("T")  # neighbours must not leak into the tail
    packed, out_len, bad = cq.pack_reads(bases, None, lengths, stride=256, threads=2)
    assert bad == 0 and np.array_equal(out_len, lengths)
    for i, r in enumerate(reads):
        want, _ = numpy_pack(r)
        assert np.array_equal(packed[i, :len(want)], want), i
        assert not packed[i, len(want):].any(), i


@pytest.mark.parametrize("L", [4, 100, 152, 252])
@pytest.mark.parametrize("threads", [1, 3])
def test_back_to_back_reads_of_one_length(L, threads):
    """Reads of one length (a multiple of 4) lying back to back take the packer's streaming path (16 reads =
    whole 64-byte blocks); invalid bytes, lower case, a read of another length in the middle and a count that
    is not a multiple of 16 must give exactly what the per-read path gives."""
    rng = np.random.default_rng(1000 + L + threads)
    n = 16 * 23 + 7
    bases = rng.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), n * L).astype(np.uint8)
    lengths = np.full(n, L, dtype=np.uint8)
    bad_reads = {0, 15, 16, 77, 200, 201, n - 1, n - 8}
    for i in sorted(bad_reads):
        bases[i * L + int(rng.integers(0, L))] = rng.choice(np.frombuffer(b"NnXR-*@\x00\xff", dtype=np.uint8))
    bases[5 * L - 1] = ord("N") if L > 4 else bases[5 * L - 1]      # last base of read 4
    if L > 4:
        bad_reads.add(4)
    short = {40: L - 1, 41: 0, 130: L - 4}                            # other lengths inside the stream
    for i, ln in short.items():
        lengths[i] = ln
    packed, out_len, bad = cq.pack_reads(bases, None, lengths, stride=L, threads=threads)
    assert packed.shape[1] == L // 4
    n_bad = 0
    for i in range(n):
        ln = int(lengths[i])
        read = bases[i * L:i * L + ln].tobytes()
        want, valid = numpy_pack(read)
        if valid:
            assert out_len[i] == ln, i
            assert np.array_equal(packed[i, :len(want)], want), i
        else:
            assert out_len[i] == 0, i
            n_bad += 1
    assert bad == n_bad and n_bad >= len(bad_reads) - 3


def test_stride_too_small_is_rejected():
    with pytest.raises(cq.CammiqError):
        cq.pack_reads(np.frombuffer(b"ACGTACGTA", dtype=np.uint8), None, np.array([9], np.uint8), stride=9,
                      packed_stride=2)


@pytest.mark.parametrize("isa", ["scalar", "avx2"])
def test_slower_isa_paths_agree(isa):
    """The dispatcher picks the widest ISA; the narrower packers are forced in a subprocess."""
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, cammiq_b200 as cq, test_pack_reads as t\n"
        "assert cq.pack_isa() in (%r, 'scalar'), cq.pack_isa()\n"
        "t.check(t.make_reads(np.random.default_rng(3), 2000, bad_every=11), 2, True)\n"
        "t.check(t.make_reads(np.random.default_rng(4), 2000, bad_every=0), 1, False)\n"
    ) % (REPO, os.path.join(REPO, "tests"), isa)
    env = dict(os.environ, CAMMIQ_PACK_ISA=isa)
    subprocess.run([sys.executable, "-c", code], check=True, env=env)
