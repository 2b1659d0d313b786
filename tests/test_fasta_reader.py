"""The library's host FASTA reader ``cls_fasta_read`` against the Python mirror of the reference's reader
(file_or_stdin.rs:76-116, sequence.rs:47-56) and, through it, against the oracle's: same records, headers and filtered
sequences on hand-written edge cases, non-ASCII texts and random files.  No GPU."""
import numpy as np
import pytest


def _check(text):
    from classeq2_b200.placement import read_fasta_native, read_fasta_text
    want = read_fasta_text(text)
    headers, bases, offsets = read_fasta_native(text)
    got = [(h, bytes(bases[int(offsets[i]):int(offsets[i + 1])]).decode()) for i, h in enumerate(headers)]
    assert got == want, (text[:200], got[:3], want[:3])
    return len(want)


def test_edge_cases(oracle):
    s = "ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTC"
    cases = [
        "", "\n", ">", ">\n", ">a", ">a\n", f">a\n{s}", f">a\n{s}\n", f">a\r\n{s}\r\n", f">a\n{s}\r",
        f"{s}\n>a\n{s}\n", f">\n{s}\n>b\n{s}\n", f">a\n{s}\n>\n>b\n{s}\n", f">a\n{s}\n>>>\n{s}\n", f">a\n{s}\n>\n{s}\n>c\n{s}\n",
        f">a\n>b\n{s}\n>c\n", f">a>b> c\n{s.lower()}\nNNNN--..{s}xyz\n\n\n{s}\n",
        f">a\n{s[:20]}\n{s[20:]}\n>b desc\n{s}\n>c\n{s[:34]}\n>d\n{s[:35]}\n", f">a\r\n\r\n{s}\r\n\r\n>b\r\n{s}",
        f">a\n{s}\n>\r\n>b\n{s}\n", f">x\r", f">a\n\r{s}\n", f">a\n >b\n{s}\n", "\r", "\r\n", ">\r\n", f">a\n{s}\n\r\n>b\r\r\n{s}\n",
        f">é日本\nACGTẗAẚCGﬅTﬆ\n>b\nｔａACGT\n",                      # non-ASCII: the four scalars whose upper case holds A / T
        f">a\nNNNN\n>b\n{s}\n", f">a\n{s}\n>b\nNNNN\n", f">a\n{s}\n>b\n",   # records whose lines hold no base at all
    ]
    total = sum(_check(t) for t in cases)
    assert total > 20
    for t in cases:                                                  # and the oracle's reader agrees with the mirror
        from classeq2_b200.placement import read_fasta_text
        assert read_fasta_text(t) == oracle.read_fasta_text(t)


@pytest.mark.parametrize("seed", range(8))
def test_random_texts(seed):
    rng = np.random.default_rng(300 + seed)
    eol = "\r\n" if seed % 4 == 1 else "\n"
    junk = list("NnRYKM-.* \t0123xyz>")
    parts = []
    for i in range(int(rng.integers(1, 300))):
        s = "".join(rng.choice(list("ACGTacgt"), int(rng.integers(0, 400))))
        for p in np.sort(rng.integers(0, len(s) + 1, int(rng.integers(0, 5))))[::-1]:
            s = s[:p] + "".join(rng.choice(junk[:-1] if p == 0 else junk, int(rng.integers(1, 4)))) + s[p:]
        width = int(rng.choice([7, 60, 80, 100000]))
        hdr = ">" * int(rng.integers(1, 3)) + (f"read_{i} x" if rng.random() < 0.95 else "")
        body = eol.join(s[a:a + width] for a in range(0, len(s), width))
        parts.append(hdr + eol + body + eol * int(rng.integers(1, 3)))
    text = "".join(parts)
    if seed % 3 == 0:
        text = text.rstrip("\r\n")
    _check(text)


def test_golden_queries_file(col_queries):
    import os
    from classeq2_b200.placement import read_fasta_native
    here = os.path.dirname(os.path.abspath(__file__))
    raw = open(os.path.join(here, "golden", "colletotrichum_queries.fasta"), "rb").read()
    headers, bases, offsets = read_fasta_native(raw)
    assert [(h, bytes(bases[int(offsets[i]):int(offsets[i + 1])]).decode()) for i, h in enumerate(headers)] == col_queries


def test_reference_alignment_file_filters_to_the_gap_free_sequences(oracle, col_queries):
    """The reference ships the 171 Colletotrichum sequences twice: gap-free, and as a MAFFT alignment (lower case, '-'
    gaps) - the input of its own build test.  The reader's filter (sequence.rs:47-56) must turn the second into the
    first: held for the oracle's reader, the Python mirror and the library's host reader."""
    import os
    from classeq2_b200.placement import read_fasta_native, read_fasta_text
    here = os.path.dirname(os.path.abspath(__file__))
    raw = open(os.path.join(here, "golden", "Colletotrichum_acutatum_gapdh_mafft.fasta"), "rb").read()
    assert b"-" in raw and raw.lower().count(b"a") > raw.count(b"A")            # aligned, mostly lower case
    want = col_queries[:171]
    assert oracle.read_fasta_text(raw.decode()) == want
    assert read_fasta_text(raw.decode()) == want
    headers, bases, offsets = read_fasta_native(raw)
    assert [(h, bytes(bases[int(offsets[i]):int(offsets[i + 1])]).decode()) for i, h in enumerate(headers)] == want


def test_byte_soup(oracle):
    """Short random texts over an alphabet that is mostly structure ('>', line ends, blanks, a few bases and non-ASCII
    letters): every combination of the record rules within a few lines."""
    import random
    from classeq2_b200.placement import read_fasta_text
    rng = random.Random(5)
    alpha = list("ACGTacgtNn>>\n\n\r -") + ["\r\n", "é", "ẗ"]
    for t in range(6000):
        s = "".join(rng.choice(alpha) for _ in range(rng.randint(0, 60)))
        _check(s)
        if t % 10 == 0:
            assert read_fasta_text(s) == oracle.read_fasta_text(s)
