"""Lock-step simulation of ``insert_hits_plain`` (classeq2_b200/csrc/kernels.cu, the opt-in ``-DCLS_INSERT_PLAIN`` variant
of the warp-owned de-duplication set and node-set histogram without shared-memory atomics) under arbitrary interleavings
of the lanes' loads and stores inside a phase (between two ``__syncwarp()``).  Checked against a dict-based model: every
distinct slot key is counted once per read, the histogram counts per node-set record are right, ``lst`` holds each bin
once.  The variant itself has not run on a GPU yet; this pins the algorithm it implements."""
import random

EMPTY = 0xFFFFFFFF


def run_read(rng, t1_size, t2_size, passes):
    t1 = [EMPTY] * t1_size; t2k = [EMPTY] * t2_size; t2c = [0] * t2_size; lst = []
    n_sets = 0
    t2_shift = 32 - (t2_size.bit_length() - 1)
    seen = set(); hist = {}
    total = 0
    for (hits, keys, sets) in passes:
        # --- reference
        fresh_ref = 0
        for l in range(32):
            if hits[l] and keys[l] not in seen:
                seen.add(keys[l]); hist[sets[l]] = hist.get(sets[l], 0) + 1; fresh_ref += 1
        # --- simulated kernel
        hm = [l for l in range(32) if hits[l]]
        if not hm:
            continue
        pending = [False] * 32; fresh = [False] * 32
        for l in hm:
            same = [x for x in hm if keys[x] == keys[l]]
            pending[l] = same[0] == l
        p1 = [(keys[l] >> 1) & (t1_size - 1) for l in range(32)]
        while any(pending):
            cur = [EMPTY] * 32
            # phase 1: per lane: load, then maybe store; interleave randomly across lanes
            events = []
            for l in range(32):
                if pending[l]:
                    events.append((l, 'R'))
            rng.shuffle(events)
            # build an interleaving: each lane's W comes after its R, at a random later position
            order = []
            for (l, _) in events:
                order.append((l, 'R'))
            # insert W events at random positions after the lane's R
            seq = list(order)
            for (l, _) in events:
                idx = next(i for i, e in enumerate(seq) if e == (l, 'R'))
                pos = rng.randint(idx + 1, len(seq))
                seq.insert(pos, (l, 'W'))
            for (l, ev) in seq:
                if ev == 'R':
                    cur[l] = t1[p1[l]]
                else:
                    if cur[l] == EMPTY:
                        t1[p1[l]] = keys[l]
            # barrier; phase 2
            for l in range(32):
                if pending[l]:
                    if cur[l] == keys[l]:
                        pending[l] = False
                    elif cur[l] == EMPTY and t1[p1[l]] == keys[l]:
                        fresh[l] = True; pending[l] = False
                    else:
                        p1[l] = (p1[l] + 1) & (t1_size - 1)
        fm = [l for l in range(32) if fresh[l]]
        assert len(fm) == fresh_ref, (len(fm), fresh_ref)
        if not fm:
            continue
        lead = [False] * 32; add = [0] * 32
        for l in fm:
            peers = [x for x in fm if sets[x] == sets[l]]
            lead[l] = peers[0] == l; add[l] = len(peers)
        p2 = [((sets[l] * 0x9E3779B1) & 0xFFFFFFFF) >> t2_shift for l in range(32)]
        while any(lead):
            cur = [EMPTY] * 32
            lanes = [l for l in range(32) if lead[l]]
            rng.shuffle(lanes)
            seq = [(l, 'R') for l in lanes]
            for l in lanes:
                idx = seq.index((l, 'R'))
                seq.insert(rng.randint(idx + 1, len(seq)), (l, 'W'))
            for (l, ev) in seq:
                if ev == 'R':
                    cur[l] = t2k[p2[l]]
                elif cur[l] == EMPTY:
                    t2k[p2[l]] = sets[l]
            won = [False] * 32
            for l in range(32):
                if lead[l]:
                    if cur[l] == sets[l]:
                        t2c[p2[l]] += add[l]; lead[l] = False
                    elif cur[l] == EMPTY and t2k[p2[l]] == sets[l]:
                        won[l] = True; lead[l] = False
                    else:
                        p2[l] = (p2[l] + 1) & (t2_size - 1)
            wm = [l for l in range(32) if won[l]]
            for l in wm:
                idx = n_sets + sum(1 for x in wm if x < l)
                while len(lst) <= idx: lst.append(None)
                lst[idx] = p2[l]; t2c[p2[l]] = add[l]
            n_sets += len(wm)
        total += len(fm)
    # final checks
    assert n_sets == len(hist) == len(lst), (n_sets, len(hist), len(lst))
    assert len(set(lst)) == len(lst)
    got = {t2k[p]: t2c[p] for p in lst}
    assert got == hist, (got, hist)
    assert total == len(seen)
    return total


def test_plain_tables_count_every_distinct_key_once():
    rng = random.Random(1)
    for trial in range(500):
        n_pass = rng.randint(1, 8)
        n_keys = rng.choice([4, 20, 100, 400])
        n_setv = rng.choice([1, 3, 10, 60])
        keyspace = [rng.getrandbits(31) for _ in range(n_keys)]
        if rng.random() < 0.3:                      # adversarial: many keys share a home slot
            keyspace = [(k & ~0x3FE) | (rng.randint(0, 3) << 1) for k in keyspace]
        keyspace = list(dict.fromkeys(keyspace))
        setmap = {k: rng.randint(1, n_setv) * 16 for k in keyspace}
        passes = []
        for _ in range(n_pass):
            hits = [rng.random() < 0.8 for _ in range(32)]
            keys = [rng.choice(keyspace) for _ in range(32)]
            passes.append((hits, keys, [setmap[k] for k in keys]))
        distinct = len({k for h, ks, _ in passes for hh, k in zip(h, ks) if hh})
        run_read(rng, 512 if distinct <= 232 else 1024, 256, passes)
