#!/usr/bin/env python
"""Times the two builders of the k-mer -> node-set map (cls_model_build on the host cores, cls_model_build_device on
cuda:0) on a synthetic model and checks that they give the same arrays.
usage: build_bench.py [n_tips=1000] [l_ref=1000] [seed=1001] [device-only]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from classeq2_b200 import synth  # noqa: E402
from classeq2_b200.model import BuiltModel  # noqa: E402


def main():
    n_tips = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    l_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1001
    tree = synth.make_tree(n_tips, seed)
    codes, lens = synth.make_refs(tree, l_ref, seed + 1)
    tflat = synth.tree_only_flat(tree, 35, 4)
    bases, offs = synth.refs_to_batch(codes, lens)
    out = {}
    device_only = len(sys.argv) > 4 and sys.argv[4] == "device-only"   # under ncu: one device build, nothing else
    for name, dev in ((("device", 0),) if device_only else (("device", 0), ("device_warm", 0), ("host", None))):
        t0 = time.perf_counter()
        bm = BuiltModel(tflat, tree.tip_node, bases, offs, device=dev)
        dt = time.perf_counter() - t0
        out[name] = bm.arrays()
        bm.close()
        n_occ = int(2 * np.maximum(lens.astype(np.int64) - 34, 0).sum())
        print(f"{name:12s} {dt * 1e3:9.1f} ms  {n_occ / dt / 1e6:8.1f} M occurrences/s  entries={len(out[name]['entry_hash'])} "
              f"sets={len(out[name]['set_off']) - 1}", flush=True)
    if device_only:
        return
    same = all(np.array_equal(out["host"][k], out["device"][k]) for k in ("entry_hash", "entry_bucket", "entry_set", "set_off"))
    print("same entries / set numbering:", same)


if __name__ == "__main__":
    main()
