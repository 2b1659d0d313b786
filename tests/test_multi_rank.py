"""N > 1 host logic on CPU: two `gloo` ranks shard a batch, each places its slice (with the CPU
oracle standing in for the GPU - this test is about sharding and gathering, not kernels), rank 0
gathers, and the result must equal the single-process outcome in input order."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_bounds():
    from classeq2_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


WORKER = r'''
import os, sys, json
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["CLS_ROOT"])
import classeq2_b200 as cq
from classeq2_b200.parallel import shard_batch, gather_results
from oracle import cpp_oracle, classeq_oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = os.path.join(os.environ["CLS_ROOT"], "tests", "golden")
z = np.load(os.path.join(g, "colletotrichum_model.npz"))
flat = cq.FlatModel(int(z["k_size"]), int(z["m_size"]), z["node_id"], z["node_kind"], z["child_off"], z["child_idx"],
                    z["entry_bucket"], z["entry_hash"], z["entry_set"], z["set_off"], z["set_node_ids"])
recs = O.read_fasta_text(open(os.path.join(g, "colletotrichum_queries.fasta")).read())
bases, offsets = cq.make_batch([s for _, s in recs])
b, o, lo, hi = shard_batch(bases, offsets, rank, world)
md = cpp_oracle.CppModel.from_flat(flat)
part = md.place_batch(b, o, n_threads=2)
local = cq.BatchResult(hi - lo)
for name, _ in cq.engine.RESULT_DTYPES:
    getattr(local, name)[:] = part[name]
full = gather_results(local, len(recs))
if rank == 0:
    want = md.place_batch(bases, offsets, n_threads=2)
    ok = all((getattr(full, n) == want[n]).all() for n, _ in cq.engine.RESULT_DTYPES)
    print(json.dumps({"ok": bool(ok), "n": len(recs), "world": world}))
else:
    assert full is None
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_shard_and_gather(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, CLS_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    assert '"ok": true' in line and f'"world": {world}' in line
