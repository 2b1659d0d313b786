"""Multi-GPU plumbing of the placement path: one process per GPU, queries sharded, index replicated.

The path shards by query with NO data-path collective (queries are independent in the reference:
core/src/use_cases/place_sequences/mod.rs:123-126).  What is left for
``torch.distributed`` is bookkeeping: which contiguous slice of the batch a rank places, and the
concatenation of the per-rank result arrays on rank 0 in input order (the reference's shared output
file).  Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import os

import numpy as np

from .engine import RESULT_DTYPES, BatchResult


def shard_bounds(n_queries: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of a batch of ``n_queries`` for ``rank`` of ``world``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_queries, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(bases: np.ndarray, offsets: np.ndarray, rank: int, world: int):
    """The (bases, offsets) sub-batch of ``rank``; offsets are rebased to start at 0."""
    lo, hi = shard_bounds(len(offsets) - 1, rank, world)
    off = offsets[lo:hi + 1]
    return bases[int(off[0]):int(off[-1])], (off - off[0]).astype(np.uint64), lo, hi


def gather_results(local: BatchResult, n_total: int, group=None, dst: int = 0) -> Optional[BatchResult]:
    """Concatenate the per-rank result arrays on ``dst`` in rank (= input) order.  Returns the full
    ``BatchResult`` on ``dst`` and ``None`` elsewhere.  Device-agnostic: tensors travel on the
    backend's device (CUDA for nccl, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    out = BatchResult(n_total) if rank == dst else None
    max_len = max(hi - lo for lo, hi in sizes)
    for name, dt in RESULT_DTYPES:
        arr = getattr(local, name)
        # unsigned 64/32-bit integers travel as their signed bit patterns
        view = arr.view({np.dtype(np.uint64): np.int64, np.dtype(np.uint32): np.int32}.get(arr.dtype, arr.dtype))
        # ragged slices: pad to the largest slice, all_gather (same call for nccl and gloo), trim
        pad = np.zeros(max_len, dtype=view.dtype)
        pad[: len(view)] = view
        t = torch.from_numpy(pad).to(dev)
        bufs = [torch.empty(max_len, dtype=t.dtype, device=dev) for _ in range(world)]
        dist.all_gather(bufs, t, group=group)
        if rank == dst:
            full = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)]).cpu().numpy().view(arr.dtype)
            getattr(out, name)[:] = full
    return out


# --------------------------------------------------------------------------------------------------
# Hash-sharded index (config 5 of BASELINE.json): the k-mer table is split over
# the GPUs by ``owner = (hash >> 61) % world`` and the query k-mers are routed to their owner with an
# all-to-all over NVLink (NCCL).  One process per GPU; every rank is "home" for its own reads and
# "owner" of one shard of the table.
# --------------------------------------------------------------------------------------------------
REPLY_BYTES = 12   # cls_probe_reply


def owner_of(hashes: np.ndarray, n_shards: int) -> np.ndarray:
    """Owner shard of every 64-bit k-mer hash: the top three bits modulo the number of shards."""
    return ((hashes.astype(np.uint64) >> np.uint64(61)) % np.uint64(n_shards)).astype(np.int64)


def exchange_plan(counts_to: np.ndarray, group=None) -> np.ndarray:
    """All-to-all of the per-owner request counts: returns ``counts_from[s]`` = how many hashes rank
    ``s`` sends to this rank.  Works on nccl (device tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t_in = torch.from_numpy(counts_to.astype(np.int64)).to(dev)
    t_out = torch.empty_like(t_in)
    dist.all_to_all_single(t_out, t_in, group=group)
    return t_out.cpu().numpy().astype(np.int64)


def exchange_segments(send_views, recv_views, group=None) -> None:
    """One all-to-all of ragged per-peer segments (lists of 1-D tensors, one per rank).  NCCL takes
    the views as they are (grouped send/recv over NVLink, no staging copy); gloo has no list
    all-to-all, so on CPU the segments are packed, exchanged with ``all_to_all_single`` and unpacked."""
    import torch
    import torch.distributed as dist

    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv_views, send_views, group=group)
        return
    packed_in = torch.cat(send_views) if send_views else None
    packed_out = torch.empty(sum(v.numel() for v in recv_views), dtype=packed_in.dtype)
    dist.all_to_all_single(packed_out, packed_in, [v.numel() for v in recv_views], [v.numel() for v in send_views], group=group)
    at = 0
    for v in recv_views:
        v.copy_(packed_out[at: at + v.numel()])
        at += v.numel()


def segment_views(buf, counts: np.ndarray, seg_cap: int = 0, item: int = 1):
    """Per-peer 1-D views of ``buf``: peer ``o`` owns ``counts[o] * item`` elements starting at
    ``o * seg_cap * item`` (fixed-capacity segments) or, with ``seg_cap == 0``, packed back to back."""
    if seg_cap:
        return [buf[o * seg_cap * item: (o * seg_cap + int(c)) * item] for o, c in enumerate(counts)]
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return [buf[int(off[o]) * item: int(off[o + 1]) * item] for o in range(len(counts))]


class PeerBuffers:
    """Inbox (hashes other ranks route to this one) and reply box (answers to this rank's own hashes) of
    one rank, allocated by the library and mapped into every peer process of the box through CUDA IPC.
    ``inbox[r]`` / ``reply[r]`` are the device pointers, valid in THIS process, of rank r's buffers."""

    def __init__(self, device: int, rank: int, world: int, seg_cap: int, group=None):
        import ctypes as C
        import torch.distributed as dist

        from . import _lib

        self.device, self.rank, self.world, self.seg_cap = device, rank, world, seg_cap
        self._own, self._opened = [], []

        def alloc(nbytes):
            ptr, h = C.c_void_p(), (C.c_uint8 * 64)()
            _lib.check(_lib.lib.cls_peer_alloc(device, int(nbytes), C.byref(ptr), h))
            self._own.append(ptr.value)
            return ptr.value, bytes(h)

        my_in, h_in = alloc(world * seg_cap * 8)
        my_rep, h_rep = alloc(world * seg_cap * REPLY_BYTES)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, (h_in, h_rep), group=group)
        self.inbox, self.reply = [0] * world, [0] * world
        for r in range(world):
            if r == rank:
                self.inbox[r], self.reply[r] = my_in, my_rep
                continue
            for name, hb in (("inbox", handles[r][0]), ("reply", handles[r][1])):
                ptr = C.c_void_p()
                hbuf = (C.c_uint8 * 64).from_buffer_copy(hb)
                _lib.check(_lib.lib.cls_peer_open(device, hbuf, C.byref(ptr)))
                self._opened.append(ptr.value)
                getattr(self, name)[r] = ptr.value

    def close(self):
        from . import _lib

        for p in self._opened:
            _lib.lib.cls_peer_close(self.device, p)
        for p in self._own:
            _lib.lib.cls_peer_free(self.device, p)
        self._opened, self._own = [], []


class ShardedPlacer:
    """Placement against a hash-sharded index, one instance per rank.

    ``transport="nccl"``: route_hashes -> all-to-all (hashes, 8 B each) -> shard_probe -> all-to-all back
    (replies, 12 B each) -> place_routed; the exchanges are ``torch.distributed`` collectives on the device
    buffers.  ``transport="p2p"``: the exchanges are FUSED into the kernels - the route kernel stores every
    hash straight into its owner's inbox and the probe kernel stores every reply straight into the asking
    GPU's reply box, both through NVLink peer pointers (CUDA IPC, :class:`PeerBuffers`); NCCL only carries
    the eight per-owner counts and one barrier per batch.  The compute stages are the library's CUDA kernels
    in both cases.  With ``world == 1`` no process group is needed."""

    def __init__(self, model, device: int, rank: int, world: int, group=None, slack: float = 1.15,
                 transport: str = "nccl", max_windows: int = 0):
        from .engine import Index

        self.rank, self.world, self.group, self.slack = rank, world, group, slack
        self.index = Index(model, device=device, shard=rank, n_shards=world)
        self.device = device
        self.transport = transport
        self.peers = None
        if transport == "p2p":
            if max_windows <= 0:
                raise ValueError("transport='p2p' needs max_windows (k-mer windows of the largest batch)")
            # a multiple of 32 slots: every rank's segment of a reply box then starts on a 128-byte line (32 x 12 B = 3 lines),
            # so that the probe kernel's 384-byte rows are whole lines on the wire (CLS_SEG_ALIGN=4: round 1's alignment)
            al = max(4, int(os.environ.get("CLS_SEG_ALIGN", "32")))
            self.peers = PeerBuffers(device, rank, world, (int(max_windows / world * slack) + 65536 + al - 1) // al * al, group)
        elif transport != "nccl":
            raise ValueError("transport must be 'nccl' or 'p2p'")

    def place(self, seqs, params=None) -> BatchResult:
        """Host buffers in, host arrays out: upload, the routed pipeline, fetch."""
        import torch

        rb = self.index.upload(seqs)
        self.place_resident(rb, params)
        res = rb.fetch(torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream)
        rb.close()
        return res

    def place_resident(self, rb, params=None) -> None:
        """The routed pipeline over a batch already resident in HBM, enqueued on torch's current
        stream (the host only waits for the per-owner counts, which size the exchange).  Results stay
        on the device (``rb.fetch``).  ``self.timing`` holds the CUDA-event time of every stage."""
        if self.transport == "p2p":
            return self._place_resident_p2p(rb, params)
        import torch

        dev = torch.device("cuda", self.device)
        st = torch.cuda.current_stream(dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        nw = rb.routed_windows()
        world = self.world
        seg_cap = (int(nw / world * self.slack) + 65536 + 3) & ~3   # multiple of 4: reply segments stay 16-byte aligned
        buf = self._buffers(nw, seg_cap, dev)
        send, slot_win, rep_in = buf["send"], buf["slot_win"], buf["rep_in"]
        ev[0].record(st)
        counts_to = rb.route_hashes(world, seg_cap, send.data_ptr(), slot_win.data_ptr(), st.cuda_stream).astype(np.int64)
        ev[1].record(st)
        counts_from = exchange_plan(counts_to, self.group) if world > 1 else counts_to.copy()
        n_recv = int(counts_from.sum())
        if buf["recv"].numel() < max(n_recv, 1):
            buf["recv"] = torch.empty(int(n_recv * 1.05) + 1, dtype=torch.int64, device=dev)
            buf["rep_out"] = torch.empty(buf["recv"].numel() * REPLY_BYTES, dtype=torch.uint8, device=dev)
        recv, rep_out = buf["recv"], buf["rep_out"]
        send_views, recv_views = segment_views(send, counts_to, seg_cap), segment_views(recv, counts_from)
        if world > 1:
            exchange_segments(send_views, recv_views, self.group)
        else:
            recv_views[0].copy_(send_views[0])
        ev[2].record(st)
        # owner side: answer everything received, in the order received
        self.index.shard_probe(recv.data_ptr(), n_recv, rep_out.data_ptr(), st.cuda_stream)
        ev[3].record(st)
        # replies travel back into a buffer laid out exactly like `send`
        back_send = segment_views(rep_out, counts_from, 0, REPLY_BYTES)
        back_recv = segment_views(rep_in, counts_to, seg_cap, REPLY_BYTES)
        if world > 1:
            exchange_segments(back_send, back_recv, self.group)
        else:
            back_recv[0].copy_(back_send[0])
        ev[4].record(st)
        rb.place_routed(rep_in.data_ptr(), slot_win.data_ptr(), world, seg_cap, params, st.cuda_stream)
        ev[5].record(st)
        self._events = ev
        self._stage_names = ["route_ms", "send_ms", "probe_ms", "reply_ms", "place_ms"]
        remote = int(counts_to.sum() - counts_to[self.rank]), int(counts_from.sum() - counts_from[self.rank])
        self._stats = dict(n_windows=nw, routed_out=remote[0], routed_in=remote[1],
                           wire_bytes_out=remote[0] * 8 + remote[1] * REPLY_BYTES,
                           wire_bytes_in=remote[1] * 8 + remote[0] * REPLY_BYTES)

    def _place_resident_p2p(self, rb, params=None) -> None:
        import torch
        import torch.distributed as dist

        dev = torch.device("cuda", self.device)
        st = torch.cuda.current_stream(dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        pb, world, me = self.peers, self.world, self.rank
        nw = rb.routed_windows()
        if ((int(nw / world * self.slack) + 65536 + 3) & ~3) > pb.seg_cap:
            raise ValueError("batch has more k-mer windows than the peer buffers were sized for (max_windows)")
        seg_cap = pb.seg_cap
        b = getattr(self, "_buf", None)
        if b is None:
            b = dict(slot_win=torch.empty(world * seg_cap, dtype=torch.int16, device=dev),
                     token=torch.zeros(1, dtype=torch.int32, device=dev))
            self._buf = b
        ev[0].record(st)
        # stage 1+2+3: hash, group by owner, and store into the owners' inboxes (my segment of each)
        seg_ptrs = [pb.inbox[o] + me * seg_cap * 8 for o in range(world)]
        counts_to = rb.route_hashes_p2p(world, seg_cap, seg_ptrs, b["slot_win"].data_ptr(), st.cuda_stream).astype(np.int64)
        ev[1].record(st)
        # the counts all-to-all doubles as "every rank's route kernel has completed": my inbox is whole
        counts_from = exchange_plan(counts_to, self.group) if world > 1 else counts_to.copy()
        ev[2].record(st)
        # stage 4+5: answer every sender's segment straight into that sender's reply box (my segment of it).  The ranks walk
        # the senders in STAGGERED order (me, me + 1, ...): with the same order on every rank all eight GPUs stored into ONE
        # GPU's reply box at a time - 900 GB/s of NVLink ingress shared by eight writers while seven links idled (175 GB/s
        # out per GPU, 17.5 of the 29 ms of a step; profiles/r2b).  CLS_PROBE_ORDER=same restores that order (A/B).
        same_order = os.environ.get("CLS_PROBE_ORDER") == "same"
        for j in range(world):
            s_rank = j if same_order else (me + j) % world
            self.index.shard_probe(pb.inbox[me] + s_rank * seg_cap * 8, int(counts_from[s_rank]),
                                   pb.reply[s_rank] + me * seg_cap * REPLY_BYTES, st.cuda_stream)
        if world > 1:
            dist.all_reduce(b["token"], group=self.group)  # barrier on the stream: every probe kernel has completed
        ev[3].record(st)
        rb.place_routed(pb.reply[me], b["slot_win"].data_ptr(), world, seg_cap, params, st.cuda_stream)
        ev[4].record(st)
        self._events = ev
        self._stage_names = ["route_ms", "counts_ms", "probe_ms", "place_ms"]
        remote = int(counts_to.sum() - counts_to[me]), int(counts_from.sum() - counts_from[me])
        self._stats = dict(n_windows=nw, routed_out=remote[0], routed_in=remote[1],
                           wire_bytes_out=remote[0] * 8 + remote[1] * REPLY_BYTES,
                           wire_bytes_in=remote[1] * 8 + remote[0] * REPLY_BYTES)

    def _buffers(self, nw: int, seg_cap: int, dev):
        import torch

        b = getattr(self, "_buf", None)
        if b is None or b["seg_cap"] != seg_cap:
            b = dict(seg_cap=seg_cap,
                     send=torch.empty(self.world * seg_cap, dtype=torch.int64, device=dev),
                     slot_win=torch.empty(self.world * seg_cap, dtype=torch.int16, device=dev),
                     rep_in=torch.empty(self.world * seg_cap * REPLY_BYTES, dtype=torch.uint8, device=dev),
                     recv=torch.empty(1, dtype=torch.int64, device=dev),
                     rep_out=torch.empty(REPLY_BYTES, dtype=torch.uint8, device=dev))
            self._buf = b
        return b

    @property
    def timing(self) -> dict:
        """Stage times (ms, CUDA events) and wire volume of the last ``place_resident``; synchronises."""
        ev = getattr(self, "_events", None)
        if ev is None:
            return {}
        ev[-1].synchronize()
        t = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(self._stage_names)}
        t["total_ms"] = ev[0].elapsed_time(ev[-1])
        t.update(self._stats)
        return t

    def close(self):
        if self.peers is not None:
            self.peers.close()
            self.peers = None


class LocalShardedPlacer:
    """The routed pipeline with ``n_shards`` table shards on ONE GPU and the two exchanges done as
    local copies: the same kernels and buffer layouts as :class:`ShardedPlacer`, no process group.
    Used by the parity tests (any ``n_shards`` on a single device)."""

    def __init__(self, model, device: int, n_shards: int, slack: float = 1.15):
        from .engine import Index

        self.n_shards, self.device, self.slack = n_shards, device, slack
        self.shards = [Index(model, device=device, shard=s, n_shards=n_shards) for s in range(n_shards)]

    def place(self, seqs, params=None) -> BatchResult:
        import torch

        dev = torch.device("cuda", self.device)
        rb = self.shards[0].upload(seqs)
        st = torch.cuda.current_stream(dev)
        nw = rb.routed_windows()
        seg_cap = (int(nw / self.n_shards * self.slack) + 65536 + 3) & ~3
        send = torch.empty(self.n_shards * seg_cap, dtype=torch.int64, device=dev)
        slot_win = torch.empty(self.n_shards * seg_cap, dtype=torch.int16, device=dev)
        counts = rb.route_hashes(self.n_shards, seg_cap, send.data_ptr(), slot_win.data_ptr(), st.cuda_stream)
        assert int(counts.sum()) == nw
        rep = torch.empty(self.n_shards * seg_cap * REPLY_BYTES, dtype=torch.uint8, device=dev)
        for o, ix in enumerate(self.shards):
            ix.shard_probe(send.data_ptr() + o * seg_cap * 8, int(counts[o]), rep.data_ptr() + o * seg_cap * REPLY_BYTES,
                           st.cuda_stream)
        rb.place_routed(rep.data_ptr(), slot_win.data_ptr(), self.n_shards, seg_cap, params, st.cuda_stream)
        res = rb.fetch(st.cuda_stream)
        self.last_counts, self.last_send = counts, (send, seg_cap)
        rb.close()
        return res
