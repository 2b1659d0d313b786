"""Models of the two de-duplication schemes of the round-2 kernels, checked against a set: the reference counts DISTINCT
k-mer hashes of a query (``kmers_map.rs:273-311``: HashSets), the kernels count distinct TABLE SLOTS of the hits.

* ``scan2_consume`` (scan2_kernels.cuh, one warp per read): a test-and-set filter indexed by hash bits; a hit whose bit
  was set already is checked against the slots of all earlier hits of the read (kept pass by pass) and against the
  other hits of its own pass.
* ``gather_kernel`` (frag_kernels.cuh, one CTA per read, warps in any order): pass A sets the filter bit of every hit and
  records bits that were set already in a second bitmap; pass B counts a hit whose bit never collided at once and sends
  the others through an exact set.

The atomics are linearisable, so an interleaving is an ORDER of the operations: the models draw random orders.  No GPU."""
import random

import pytest


def _random_read(rng, n_windows, n_slots, hit_rate, filter_bits):
    """(slot or None per window, filter bit per window): slots repeat (low-complexity reads), different slots share bits."""
    slots = [rng.randrange(n_slots) if rng.random() < hit_rate else None for _ in range(n_windows)]
    bit_of = {}
    return slots, [None if s is None else bit_of.setdefault(s, rng.randrange(filter_bits)) for s in slots]


@pytest.mark.parametrize("seed", range(40))
def test_scan2_filter_with_exact_check_counts_distinct_slots(seed):
    rng = random.Random(seed)
    n_windows = rng.choice([2, 33, 116, 232, 256])
    slots, bits = _random_read(rng, n_windows, rng.choice([3, 40, 1000]), rng.random(), rng.choice([8, 64, 8192]))
    filt, earlier, fresh_total = set(), [], 0
    for p0 in range(0, n_windows, 32):                           # one pass = 32 lanes
        lanes = list(range(p0, min(p0 + 32, n_windows)))
        order = lanes[:]
        rng.shuffle(order)                                       # the order in which the lanes' ATOMS.OR land
        candidate = {}
        for l in order:
            if slots[l] is None:
                continue
            candidate[l] = bits[l] in filt
            filt.add(bits[l])
        fresh = {l: slots[l] is not None for l in lanes}
        cand = [l for l in lanes if candidate.get(l)]            # ascending lane order, as the ballot is walked
        for l in cand:
            same = [x for x in lanes if x != l and slots[x] == slots[l]]
            dup = slots[l] in earlier or any(x not in cand for x in same) or any(x < l for x in same)
            if dup:
                fresh[l] = False
        fresh_total += sum(fresh.values())
        earlier += [slots[l] for l in lanes if slots[l] is not None]
    assert fresh_total == len({s for s in slots if s is not None})


@pytest.mark.parametrize("seed", range(40))
def test_gather_two_bitmaps_and_exact_set_count_distinct_slots(seed):
    rng = random.Random(1000 + seed)
    n_windows = rng.choice([5, 600, 3032])
    slots, bits = _random_read(rng, n_windows, rng.choice([2, 50, 5000]), rng.random(), rng.choice([16, 1024, 65536]))
    hits = [i for i in range(n_windows) if slots[i] is not None]
    # pass A: every hit test-and-sets its bit, in any order; a bit found set is recorded as collided
    order = hits[:]
    rng.shuffle(order)
    filt, collided = set(), set()
    for i in order:
        if bits[i] in filt:
            collided.add(bits[i])
        filt.add(bits[i])
    # (barrier) pass B: a hit of a bit that never collided is distinct; the others insert their slot into an exact set
    order = hits[:]
    rng.shuffle(order)
    exact, fresh = set(), 0
    for i in order:
        if bits[i] not in collided:
            fresh += 1
        elif slots[i] not in exact:
            exact.add(slots[i])
            fresh += 1
    assert fresh == len({slots[i] for i in hits})
