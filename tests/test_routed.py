"""Hash-sharded index (config 5).

GPU tests: the routed pipeline (route_hashes -> probe per shard -> place_routed) with 1, 2, 3 and 8
table shards on ONE device must give, field for field, the results of the replicated index; every
hash must sit in the segment of its owner; shard tables must partition the entries.
CPU tests: two/three `gloo` ranks run the same exchange plumbing (counts all-to-all, ragged segment
all-to-all, replies back into the send layout) with numpy stand-ins for the three kernels."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _synthetic(n_tips=300, l_ref=400, n_reads=3000, seed=77, read_len=150):
    from classeq2_b200 import synth
    sm = synth.make_model(n_tips, l_ref, seed)
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, n_reads, read_len, seed + 2)
    return sm, bases, offsets


def _equal(a, b):
    from classeq2_b200.engine import RESULT_DTYPES
    return {n: int((getattr(a, n) != getattr(b, n)).sum()) for n, _ in RESULT_DTYPES}


@pytest.mark.gpu
@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_routed_equals_replicated_synthetic(n_shards):
    import classeq2_b200 as cq
    from classeq2_b200.parallel import LocalShardedPlacer
    sm, bases, offsets = _synthetic()
    want = cq.Index(sm.flat, device=0).place_batch((bases, offsets))
    lp = LocalShardedPlacer(sm.flat, 0, n_shards)
    got = lp.place((bases, offsets))
    assert all(v == 0 for v in _equal(got, want).values()), _equal(got, want)
    # shard tables partition the index entries
    total = cq.Index(sm.flat, device=0).info()["n_entries"]
    assert sum(ix.info()["n_entries"] for ix in lp.shards) == total
    # every routed hash sits in its owner's segment, and the multiset of routed hashes is the batch's
    from classeq2_b200.parallel import owner_of
    send, seg_cap = lp.last_send
    h = send.cpu().numpy().view(np.uint64)
    for o in range(n_shards):
        seg = h[o * seg_cap: o * seg_cap + int(lp.last_counts[o])]
        assert (owner_of(seg, n_shards) == o).all()
    from oracle import cpp_oracle
    ref = np.concatenate([cpp_oracle.kmer_hashes(bases[int(offsets[i]):int(offsets[i + 1])].tobytes(), 35) for i in range(50)])
    routed = np.concatenate([h[o * seg_cap: o * seg_cap + int(lp.last_counts[o])] for o in range(n_shards)])
    assert np.isin(ref, routed).all()
    assert int(lp.last_counts.sum()) == int((2 * (np.diff(offsets.astype(np.int64)) - 34)).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("general", [False, True])
def test_routed_colletotrichum_fixture(col_flat, col_queries, col_expected, general):
    """The reference's own data fixture (multifurcating tree, queries of up to 160 bases and unrelated
    negatives) through four shards, closed and general node-set records."""
    import classeq2_b200 as cq
    from classeq2_b200.parallel import LocalShardedPlacer
    seqs = [s for _, s in col_queries if len(s) <= 160]
    seqs += [s[a:a + 150] for _, s in col_queries if len(s) > 160 for a in (0, 40, len(s) - 150)]  # 150-base windows of the long ones
    flat = col_flat.with_general_sets() if general else col_flat
    want = cq.Index(flat, device=0).place_batch(seqs)
    got = LocalShardedPlacer(flat, 0, 4).place(seqs)
    assert all(v == 0 for v in _equal(got, want).values()), _equal(got, want)
    for p in (cq.PlaceParams(remove_intersection=True), cq.PlaceParams(min_match_coverage=0.2, max_iterations=3)):
        want = cq.Index(flat, device=0).place_batch(seqs, p)
        got = LocalShardedPlacer(flat, 0, 4).place(seqs, p)
        assert all(v == 0 for v in _equal(got, want).values()), _equal(got, want)


@pytest.mark.gpu
def test_routed_ragged_and_edge_batches():
    import classeq2_b200 as cq
    from classeq2_b200.parallel import LocalShardedPlacer
    sm, bases, offsets = _synthetic(n_tips=120, l_ref=300, n_reads=0)
    rng = np.random.default_rng(5)
    lens = rng.integers(20, 162, 700)   # some shorter than k: decided on the host
    from classeq2_b200 import synth
    b2, o2, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, len(lens), lens, 9)
    want = cq.Index(sm.flat, device=0).place_batch((b2, o2))
    got = LocalShardedPlacer(sm.flat, 0, 8).place((b2, o2))
    assert all(v == 0 for v in _equal(got, want).values()), _equal(got, want)
    empty = LocalShardedPlacer(sm.flat, 0, 2).place([])
    assert empty.n == 0
    # a sharded handle refuses the replicated calls, and reads beyond the warp-per-read geometry are refused
    ix = cq.Index(sm.flat, device=0, shard=1, n_shards=2)
    with pytest.raises(cq._lib.ClsError):
        ix.place_batch(["ACGT" * 20])
    with pytest.raises(cq._lib.ClsError):
        LocalShardedPlacer(sm.flat, 0, 2).place(["ACGT" * 400])
    with pytest.raises(cq._lib.ClsError):
        cq.Index(sm.flat, device=0, shard=2, n_shards=2)


# ---- CPU: the exchange plumbing under gloo -----------------------------------------------------
def test_owner_of_and_segment_views():
    from classeq2_b200.parallel import owner_of, segment_views
    h = np.array([0, 1 << 61, 3 << 61, (7 << 61) + 5, (1 << 64) - 1], dtype=np.uint64)
    assert owner_of(h, 8).tolist() == [0, 1, 3, 7, 7]
    assert owner_of(h, 2).tolist() == [0, 1, 1, 1, 1]
    assert owner_of(h, 3).tolist() == [0, 1, 0, 1, 1]
    assert owner_of(h, 1).tolist() == [0, 0, 0, 0, 0]
    buf = np.arange(40)
    v = segment_views(buf, np.array([2, 0, 3]), seg_cap=10)
    assert [x.tolist() for x in v] == [[0, 1], [], [20, 21, 22]]
    v = segment_views(buf, np.array([2, 0, 3]), seg_cap=0, item=2)
    assert [x.tolist() for x in v] == [[0, 1, 2, 3], [], [4, 5, 6, 7, 8, 9]]


WORKER = r'''
import os, sys, json
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["CLS_ROOT"])
from classeq2_b200.parallel import owner_of, exchange_plan, exchange_segments, segment_views
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# every rank "hashes" its own reads: deterministic pseudo-hashes; the "index" holds the multiples of 3
rng = np.random.Generator(np.random.PCG64(100 + rank))
hashes = rng.integers(0, 2**64, 5000 + 100 * rank, dtype=np.uint64)
own = owner_of(hashes, world)
seg_cap = len(hashes)
send = np.zeros(world * seg_cap, dtype=np.uint64)
win_slot = np.zeros(len(hashes), dtype=np.int64)
counts_to = np.zeros(world, dtype=np.int64)
for i, (h, o) in enumerate(zip(hashes, own)):      # stand-in for cls_route_hashes
    win_slot[i] = o * seg_cap + counts_to[o]
    send[win_slot[i]] = h
    counts_to[o] += 1
counts_from = exchange_plan(counts_to)
t_send = torch.from_numpy(send.view(np.int64))
recv = torch.empty(int(counts_from.sum()), dtype=torch.int64)
exchange_segments(segment_views(t_send, counts_to, seg_cap), segment_views(recv, counts_from))
got = recv.numpy().view(np.uint64)
assert (owner_of(got, world) == rank).all()      # only hashes this rank owns arrive here
rep_out = np.where(got % np.uint64(3) == 0, got // np.uint64(3), np.uint64(0xFFFFFFFF)).astype(np.uint64)  # stand-in for cls_shard_probe
rep_in = torch.zeros(world * seg_cap, dtype=torch.int64)
exchange_segments(segment_views(torch.from_numpy(rep_out.view(np.int64)), counts_from),
                  segment_views(rep_in, counts_to, seg_cap))
ans = rep_in.numpy().view(np.uint64)[win_slot]  # stand-in for cls_place_routed's reply lookup
want = np.where(hashes % np.uint64(3) == 0, hashes // np.uint64(3), np.uint64(0xFFFFFFFF))
ok = bool((ans == want).all())
flags = [None] * world
dist.all_gather_object(flags, ok)
if rank == 0:
    print(json.dumps({"ok": all(flags), "world": world, "routed": int(counts_to.sum())}))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_routed_exchange(tmp_path, world):
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
    assert p.returncode == 0, p.stderr[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    assert '"ok": true' in line and f'"world": {world}' in line


NCCL_WORKER = r'''
import os, sys, json
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["CLS_ROOT"])
import classeq2_b200 as cq
from classeq2_b200 import synth
from classeq2_b200.parallel import ShardedPlacer
from classeq2_b200.engine import RESULT_DTYPES
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sm = synth.make_model(300, 400, 77)
bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, 4000 + 500 * rank, 150, 79 + rank)   # every rank places its OWN reads
want = cq.Index(sm.flat, device=local).place_batch((bases, offsets))
sp = ShardedPlacer(sm.flat, local, rank, world)
got = sp.place((bases, offsets))
bad = {n: int((getattr(got, n) != getattr(want, n)).sum()) for n, _ in RESULT_DTYPES}
# the same with the exchange fused into the kernels (peer stores over NVLink instead of NCCL all-to-alls), twice
# in a row so that the reuse of the inboxes / reply boxes is exercised
sp2 = ShardedPlacer(sm.flat, local, rank, world, transport="p2p", max_windows=2 * 116 * (4000 + 500 * world))
for _ in range(2):
    got2 = sp2.place((bases, offsets))
    for n, _ in RESULT_DTYPES:
        bad[n] += int((getattr(got2, n) != getattr(want, n)).sum())
sp2.close()
flags = [None] * world
dist.all_gather_object(flags, (bad, sp.timing["routed_out"], sp.index.info()["n_entries"]))
if rank == 0:
    print(json.dumps({"ok": all(not any(f[0].values()) for f in flags), "world": world, "routed_out": [f[1] for f in flags],
                      "entries": [f[2] for f in flags], "total_entries": cq.Index(sm.flat, device=local).info()["n_entries"]}))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_nccl_sharded_placer_all_gpus(tmp_path):
    """Every visible GPU holds one shard; k-mers cross NVLink through NCCL all-to-alls; the placements
    must equal the replicated index's, rank by rank.  Needs at least two GPUs."""
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(NCCL_WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, CLS_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    import json
    out = json.loads(line)
    assert out["ok"] and out["world"] == world
    assert sum(out["entries"]) == out["total_entries"] and all(r > 0 for r in out["routed_out"])
