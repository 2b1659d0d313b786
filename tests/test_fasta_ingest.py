"""FASTA ingest on the device (cls_fasta_upload) against the host mirror of the reference's reader
(file_or_stdin.rs:76-116, sequence.rs:47-56): same records, same headers, same filtered sequences -
checked through the placements and k-mer counts they produce - on hand-written edge cases and on
random texts that cross the 4 KiB tiles of the scan in every possible way."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import classeq2_b200 as cq
    from classeq2_b200 import synth
    sm = synth.make_model(80, 400, 55)
    return cq, sm, cq.Index(sm.flat, device=0)


def _check(cq, ix, text: str):
    from classeq2_b200.placement import read_fasta_text
    want = read_fasta_text(text)
    rb, headers, lengths = ix.upload_fasta(text.encode())
    assert headers == [h for h, _ in want], (text[:200], headers[:5], want[:5])
    assert lengths.tolist() == [len(s) for _, s in want]
    rb.place()
    got = rb.fetch()
    ref = ix.place_batch([s for _, s in want])
    for f, _ in cq.engine.RESULT_DTYPES:
        assert (getattr(got, f) == getattr(ref, f)).all(), (f, text[:200])
    rb.close()
    return len(want)


def test_edge_cases(setup):
    cq, sm, ix = setup
    s = "ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTC"
    cases = [
        "", "\n", ">", ">\n", ">a", ">a\n", f">a\n{s}", f">a\n{s}\n", f">a\r\n{s}\r\n", f">a\n{s}\r",
        f"{s}\n>a\n{s}\n",                               # sequence before any header: nothing is sent
        f">\n{s}\n>b\n{s}\n",                            # empty header with sequence: the reader stops at >b
        f">a\n{s}\n>\n>b\n{s}\n",                        # empty header without sequence: harmless
        f">a\n{s}\n>>>\n{s}\n",                          # trailing empty header: its sequence is never sent
        f">a\n{s}\n>\n{s}\n>c\n{s}\n",                   # ... mid-file: stops there, >a was sent
        f">a\n>b\n{s}\n>c\n",                            # empty body mid-file is sent, trailing one is dropped
        f">a>b> c\n{s.lower()}\nNNNN--..{s}xyz\n\n\n{s}\n",   # '>' removed everywhere, filtering, blank lines
        f">a\n{s[:20]}\n{s[20:]}\n>b desc\n{s}\n>c\n{s[:34]}\n>d\n{s[:35]}\n",   # wrapped lines; too short / just long enough
        f">a\r\n\r\n{s}\r\n\r\n>b\r\n{s}",              # CRLF files with blank lines
        f">a\n{s}\n>\r\n>b\n{s}\n", f">x\r",            # "\r\n" terminator after '>' (empty header); "\r" at EOF stays in the header
        f">a\n\r{s}\n", f">a\n >b\n{s}\n",              # a line starting with "\r" or a blank is a sequence line
    ]
    total = sum(_check(cq, ix, t) for t in cases)
    assert total > 15


def test_non_ascii_is_refused(setup):
    cq, sm, ix = setup
    with pytest.raises(cq._lib.ClsError) as e:
        ix.upload_fasta(">a\nACGTẗACGT\n".encode())
    assert e.value.code == cq._lib.CLS_ERR_UNSUPPORTED


@pytest.mark.parametrize("seed", range(12))
def test_random_texts(setup, seed):
    cq, sm, ix = setup
    from classeq2_b200 import synth
    rng = np.random.default_rng(900 + seed)
    n_reads = int(rng.integers(1, 400))
    lens = rng.integers(20, 700 if seed % 3 else 3500, n_reads)     # 4.1 kb is the longest read the one-CTA geometry holds
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, n_reads, lens, 77 + seed)
    eol = "\r\n" if seed % 4 == 1 else "\n"
    width = int(rng.choice([60, 70, 80, 150, 100000]))
    junk = list("NnRYKM-.* \t0123xyz")
    parts = []
    for i in range(n_reads):
        s = bases[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()
        if rng.random() < 0.5:
            s = s.lower() if rng.random() < 0.5 else "".join(c.lower() if rng.random() < 0.3 else c for c in s)
        if rng.random() < 0.3:      # sprinkle characters the filter deletes
            pos = np.sort(rng.integers(0, len(s) + 1, int(rng.integers(1, 6))))[::-1]
            for p in pos:
                s = s[:p] + "".join(rng.choice(junk, int(rng.integers(1, 4)))) + s[p:]
        if rng.random() < 0.05:     # a run of deleted characters longer than two 4 KiB tiles: tiles without any line start
            p = int(rng.integers(0, len(s) + 1))
            s = s[:p] + "N" * int(rng.integers(8200, 13000)) + s[p:]
        hdr = f">read_{i} len={len(s)}" + (" >x" if rng.random() < 0.1 else "") + (" pad" * 3000 if rng.random() < 0.02 else "")
        body = eol.join(s[a:a + width] for a in range(0, len(s), width))
        parts.append(hdr + eol + body + (eol if rng.random() < 0.95 else eol + eol))
        if rng.random() < 0.03:
            parts.append(">empty_mid" + eol)
    text = "".join(parts)
    if seed % 5 == 0:
        text = text.rstrip("\r\n")            # no terminator after the last line
    if seed == 7:
        text = "\n\n" + text                  # leading blank lines
    assert _check(cq, ix, text) >= n_reads
