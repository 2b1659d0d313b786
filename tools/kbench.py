#!/usr/bin/env python
"""Quick kernel-resident timing of the placement kernel for A/B experiments (not the bench).
usage: [CLASSEQ_B200_LIB=...] python tools/kbench.py [config] [n_reads] [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import classeq2_b200 as cq  # noqa: E402
from classeq2_b200 import synth  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
c = synth.CONFIGS[cfg]
cache = f"/tmp/kbench_{cfg}_{n}.npz"
sm = synth.make_model(c["n_tips"], c["l_ref"], c["tree_seed"])
lens = synth.skewed_lengths(n, c["len_seed"]) if c["read_len"] == "skewed" else c["read_len"]
bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, n, lens, c["tree_seed"] + 2)
ix = cq.Index(sm.flat, device=0)
rb = ix.upload((bases, offsets))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
ts = []
with torch.cuda.stream(st):
    for i in range(steps + 2):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        rb.place(None, st.cuda_stream)
        b.record(st)
        ts.append((a, b))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ts][2:]
res = rb.fetch(st.cuda_stream)
print(f"{os.path.basename(cq._lib.LIB_PATH)} config{cfg} n={n}: {np.mean(ms):.3f} ms/step (min {min(ms):.3f}) "
      f"{n / np.mean(ms) / 1e3:.1f} M reads/s  status_hist={np.bincount(res.status, minlength=11).tolist()} "
      f"sum(node)={int(res.node_id.sum())} sum(one)={int(res.one.sum())} info={ix.info()['closed_sets']}")
