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


def _read_fasta_bytes(raw: bytes):
    """The reader on BYTES: ``BufRead::lines`` yields ``Err`` for a line that is not valid UTF-8 and the reader returns
    there (`line?`, file_or_stdin.rs:85) - the records sent so far stand, the pending one is never sent.  Up to that
    line this is read_fasta_text's state machine."""
    from classeq2_b200.placement import filter_sequence
    out, header, parts = [], "", []
    lines = raw.split(b"\n")
    terminated = [True] * len(lines)
    terminated[-1] = False
    if lines and lines[-1] == b"":
        lines.pop()
    for bline, term in zip(lines, terminated):
        if term and bline.endswith(b"\r"):
            bline = bline[:-1]
        if bline == b"":
            continue
        try:
            line = bline.decode("utf-8")
        except UnicodeDecodeError:
            return out
        if line.startswith(">"):
            if header != "":
                out.append((header, "".join(parts)))
                parts = []
            elif parts:
                return out
            header = line.replace(">", "")
        else:
            f = filter_sequence(line)
            if f:
                parts.append(f)
    if header != "" and parts:
        out.append((header, "".join(parts)))
    return out


def _native_records(raw: bytes):
    from classeq2_b200.placement import read_fasta_native
    headers, bases, offsets = read_fasta_native(raw)
    return [(h, bytes(bases[int(offsets[i]):int(offsets[i + 1])]).decode()) for i, h in enumerate(headers)]


def test_reader_stops_at_a_line_that_is_not_valid_utf8():
    s = "ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTC"
    good = f">a\n{s}\n>b\n{s}\n".encode()
    cases = [
        (good + b">c\n" + s.encode() + b"\n\xff\n>d\n" + s.encode() + b"\n", 2),          # the pending record c is not sent
        (good + b">c \xc3\n" + s.encode() + b"\n", 1),                                    # truncated sequence in a header: b is pending, not sent
        (good + b"\xed\xa0\x80\n", 1),                                                    # a surrogate
        (good + b"\xc0\xaf\n", 1),                                                        # an overlong form
        (good + b"\xf4\x90\x80\x80\n", 1),                                                # beyond U+10FFFF
        (b"\x80\n" + good, 0),
        (good + ">é日\n".encode() + s.encode() + b"\n", 3),                               # valid non-ASCII goes on
    ]
    for raw, n in cases:
        got = _native_records(raw)
        assert got == _read_fasta_bytes(raw) and len(got) == n, (raw[-40:], len(got), n)


@pytest.mark.parametrize("chunk", ["1", "3", "17", "64", "1000"])
def test_chunked_reader_gives_the_same_records_whatever_the_chunk_size(monkeypatch, chunk):
    """cls_fasta_read scans chunks of whole lines on the host pool and runs the reader's state machine over their events
    (header lines, invalid lines): with chunks of a few bytes (CLS_FASTA_CHUNK) every record rule meets a chunk boundary -
    header at the end of a chunk, wrapped sequences over several, the stop conditions inside a later chunk of a wave and
    in a later wave."""
    import random
    monkeypatch.setenv("CLS_FASTA_CHUNK", chunk)
    monkeypatch.setenv("CLS_HOST_THREADS", "4")
    rng = random.Random(int(chunk))
    s = "ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTC"
    fixed = [f">a\n{s}\n>b\n{s[:30]}\n{s[30:]}\n>c\r\n{s}\r\n", f"{s}\n>a\n{s}\n", f">a\n{s}\n>\n{s}\n>c\n{s}\n", f">a\n>b\n{s}\n>c\n",
             f">a\n{s}\n>\n>b\n{s}\n", "", "\n\n\n", ">a", f">a\n{s}"] + [f">r{i}\n{s}\n" * 40 for i in range(2)]
    for t in fixed:
        assert _native_records(t.encode()) == _read_fasta_bytes(t.encode()), t[:80]
    alpha = [b"A", b"C", b"G", b"T", b"a", b"n", b">", b">", b"\n", b"\n", b"\n", b"\r", b" ", b"-", b"\r\n", "é".encode(), "ẗ".encode()]
    for t in range(1500):
        parts = [rng.choice(alpha) for _ in range(rng.randint(0, 120))]
        if t % 3 == 0 and parts:                                  # one byte that breaks UTF-8 somewhere
            parts[rng.randrange(len(parts))] = bytes([rng.choice([0xFF, 0x80, 0xC3, 0xE1])])
        raw = b"".join(parts)
        assert _native_records(raw) == _read_fasta_bytes(raw), raw
