import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import classeq2_b200 as cq
from classeq2_b200 import synth
c = synth.CONFIGS[2]
sm = synth.make_model(c["n_tips"], c["l_ref"], c["tree_seed"])
bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, 1000000, 150, c["tree_seed"] + 2)
ix = cq.Index(sm.flat, device=0)
out = cq.BatchResult(1000000)
for _ in range(3): ix.place_batch_into(bases, offsets, out)
t0 = time.perf_counter()
for _ in range(10): ix.place_batch_into(bases, offsets, out)
dt = (time.perf_counter() - t0) / 10
print(os.environ.get("CLS_CHUNK_MBASES"), f"e2e {dt*1e3:.2f} ms  {1e6/dt/1e6:.1f} M reads/s", {k: round(v, 2) for k, v in ix.timing().items()})
