"""A word-level model of the fragment decomposition of kb-scale reads (classeq2_b200/csrc/frag_kernels.cuh): the windows of
the fragments tile the windows of both strands exactly once, and ``decode_frag16``'s arithmetic on the 2-bit packed
words (aligned forward words; the reverse-complement stretch cut out of the read at ANY base offset with funnel shifts,
reversed word by word) yields the bases it should.  Pure Python / numpy: no GPU."""
import numpy as np
import pytest

FRAG_WINDOWS, K = 128, 35
FRAG_BASES = FRAG_WINDOWS + K - 1          # 162
CODE = {"A": 0, "C": 1, "T": 2, "G": 3}    # (ascii >> 1) & 3; complement = code ^ 2
LETTER = "ACTG"
M32 = 0xFFFFFFFF


def pack(seq):
    words = [0] * ((len(seq) + 15) // 16 + 2)   # slack words, as the packed buffers have
    for i, ch in enumerate(seq):
        words[i // 16] |= CODE[ch] << (2 * (i % 16))
    return words


def funnelshift_r(lo, hi, sh):
    sh &= 31
    return (((hi << 32) | lo) >> sh) & M32


def revcomp16(w):
    r = int(f"{w:032b}"[::-1], 2)                       # __brev
    r = ((r >> 1) & 0x55555555) | ((r & 0x55555555) << 1)
    return (r ^ 0xAAAAAAAA) & M32


def decode_frag16(words, s_f, s_r, flen):
    """The 32 lanes of decode_frag16: returns (forward bases, reverse-complement bases) of the fragment."""
    nw = (flen + 15) >> 4
    pad2 = 2 * (nw * 16 - flen)
    fw, u = [0] * 16, [0] * 16
    for t in range(16):
        if t < nw:
            fw[t] = words[(s_f >> 4) + t]
            q = (s_r >> 4) + t
            u[t] = funnelshift_r(words[q], words[q + 1], 2 * (s_r & 15))
    r = [revcomp16(x) for x in u]
    v_f, v_r = fw, [0] * 16
    for t in range(16):
        a, b = r[(nw - 1 - t) & 15], r[(nw - 2 - t) & 15]
        v_r[t] = funnelshift_r(a, b if t + 1 < nw else 0, pad2) if t < nw else 0
    dec = lambda ws: "".join(LETTER[(ws[i // 16] >> (2 * (i % 16))) & 3] for i in range(flen))  # noqa: E731
    return dec(v_f), dec(v_r)


def revcomp(s):
    return "".join({"A": "T", "C": "G", "G": "C", "T": "A"}[c] for c in reversed(s))


@pytest.mark.parametrize("L", [35, 36, 161, 162, 163, 290, 291, 300, 400, 1024, 1550, 1600, 4129])
def test_fragments_tile_the_windows_and_decode_to_the_right_bases(L):
    rng = np.random.default_rng(L)
    read = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, L))
    rc = revcomp(read)
    words = pack(read)
    W = L - K + 1
    frags = (W + FRAG_WINDOWS - 1) // FRAG_WINDOWS
    seen = np.zeros(2 * W, np.int32)
    for f in range(frags + 2):                         # units past a read's last fragment are skipped
        w0 = f * FRAG_WINDOWS
        if w0 >= W:
            continue
        flen = min(FRAG_BASES, L - w0)
        fwd, rev = decode_frag16(words, w0, L - w0 - flen, flen)
        assert fwd == read[w0:w0 + flen]
        assert rev == rc[w0:w0 + flen]                 # the same stretch of the reverse-complement strand
        wf = flen - K + 1
        assert 1 <= wf <= FRAG_WINDOWS
        for strand in (0, 1):
            seen[strand * W + w0: strand * W + w0 + wf] += 1
    assert (seen == 1).all()
