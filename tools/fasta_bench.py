#!/usr/bin/env python
"""Times cls_fasta_upload (FASTA text -> packed resident batch, on the device) on a synthetic file of
config 2's reads.  usage: python tools/fasta_bench.py [n_reads]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import classeq2_b200 as cq  # noqa: E402
from classeq2_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
c = synth.CONFIGS[2]
sm = synth.make_model(c["n_tips"], c["l_ref"], c["tree_seed"])
bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, n, 150, c["tree_seed"] + 2)
# one-line FASTA records: ">r<i>\n<150 bases>\n"
hdr = np.char.add(np.char.add(">r", np.arange(n).astype(str)), "\n").astype("S")
body = bases.reshape(n, 150)
parts = []
t0 = time.time()
buf = bytearray()
for i in range(n):
    buf += hdr[i]
    buf += body[i].tobytes()
    buf += b"\n"
text = np.frombuffer(bytes(buf), dtype=np.uint8)
print(f"text: {text.size / 1e6:.1f} MB built in {time.time() - t0:.1f} s")
ix = cq.Index(sm.flat, device=0)
for rep in range(4):
    t0 = time.perf_counter()
    rb, headers_lens = None, None
    import ctypes as C
    from classeq2_b200 import _lib
    h, rec = C.c_void_p(), _lib.FastaRecords()
    _lib.check(_lib.lib.cls_fasta_upload(ix._h, text.ctypes.data_as(_lib.u8p), text.size, C.byref(h), C.byref(rec)))
    dt = time.perf_counter() - t0
    print(f"cls_fasta_upload: {dt * 1e3:.1f} ms for {rec.n_records} records = {rec.n_records / dt / 1e6:.1f} M reads/s, "
          f"{text.size / dt / 1e9:.2f} GB/s of text")
    rbo = cq.ResidentBatch._from_handle(ix, h, int(rec.n_records))
    if rep == 3:
        rbo.place()
        got = rbo.fetch()
        want = ix.place_batch((bases, offsets))
        print("placements equal to the ASCII-batch path:", all((getattr(got, f) == getattr(want, f)).all() for f, _ in cq.engine.RESULT_DTYPES))
    rbo.close()
