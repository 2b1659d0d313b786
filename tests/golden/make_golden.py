#!/usr/bin/env python
"""Regenerates the committed fixtures under tests/golden/ (run in the build container only).

Inputs come from the reference checkout's own DATA fixtures (read-only, /root/reference):

  core/src/tests/data/colletotrichum-acutatom-complex/inputs/Colletotrichum_acutatum_gapdh-PhyML.nwk
  core/src/tests/data/colletotrichum-acutatom-complex/inputs/Colletotrichum_acutatum_gapdh_gapsfree.fasta
  core/src/tests/data/colletotrichum-acutatom-complex/outputs/Colletotrichum_acutatum_gapdh-PhyML.yaml  (ids only)
  tests/data/public/019051d9-4c7a-7b2d-9dd1-66ef92236fd7/input/*.fasta   (gyrB queries, negatives)
  tests/data/public/019051d9-4c7a-7b2d-9dd1-66ef92236fd7/output/result.yaml (wire shape + `one: 3754`)
  core/src/use_cases/place_sequences/place_sequence.rs:620-623 (the hard-coded test query)

and everything derived is computed by the CPU oracle (oracle/classeq_oracle.py).  The reference is
Rust and cannot run here, so the expected placements below are ORACLE outputs ("parity unpinned"
for placement decisions, see the oracle's header); the reference-pinned facts are kept apart in
``reference_pins.json``.

Outputs (all small):
  colletotrichum_model.npz      flat model arrays (cls_model_view layout) + tree JSON
  colletotrichum_queries.fasta  queries: 171 tips, test query, 14 gyrB negatives, seeded sub-reads/mutants
  colletotrichum_expected.json  oracle outcome per query for several knob settings
  reference_pins.json           facts pinned by the reference's own docs / fixtures
"""
import hashlib
import json
import os
import re
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import classeq_oracle as O  # noqa: E402

REF = "/root/reference"
COL = f"{REF}/core/src/tests/data/colletotrichum-acutatom-complex"
FD7 = f"{REF}/tests/data/public/019051d9-4c7a-7b2d-9dd1-66ef92236fd7"

TEST_QUERY = ("Col_orchidophilum",
              "CCTTCATTGAGACCAAGTACGCTGTGAGTATCACCCCACTTTACCCCTCCATGATGATATCACATCTGTCACGACAATACCAGCCTCATCGGCC"
              "ACTGGGAAAGAAATGAGCTAGCACTCTCGATCCTGTGACCCAGGATACTGAAGCGGCTCGTCCCAATGGCATGATGTGA")

KNOBS = [
    dict(name="default", max_iterations=None, min_match_coverage=None, remove_intersection=None),
    dict(name="remove_intersection", max_iterations=None, min_match_coverage=None, remove_intersection=True),
    dict(name="cov1", max_iterations=None, min_match_coverage=1.5, remove_intersection=False),
    dict(name="iter2", max_iterations=2, min_match_coverage=0.0, remove_intersection=False),
]


def flatten(tree: O.Tree):
    clades = list(tree.root.walk())
    idx = {id(c): i for i, c in enumerate(clades)}
    kind = {"ROOT": 0, "NODE": 1, "LEAF": 2}
    node_id = np.array([c.id for c in clades], np.uint64)
    node_kind = np.array([kind[c.kind] for c in clades], np.uint8)
    child_off = np.zeros(len(clades) + 1, np.uint64)
    child_idx = []
    for i, c in enumerate(clades):
        child_idx += [idx[id(ch)] for ch in (c.children or [])]
        child_off[i + 1] = len(child_idx)
    km = tree.kmers_map
    eb, eh, es, set_off, set_nodes, seen = [], [], [], [0], [], {}
    for key in sorted(km.map):
        for h in sorted(km.map[key]):
            fs = frozenset(km.map[key][h])
            s = seen.get(fs)
            if s is None:
                s = seen[fs] = len(seen)
                set_nodes += sorted(fs)
                set_off.append(len(set_nodes))
            eb.append(key), eh.append(h), es.append(s)
    u = lambda a: np.array(a, np.uint64)  # noqa: E731
    return dict(k_size=np.uint32(km.k_size), m_size=np.uint32(km.m_size), node_id=node_id, node_kind=node_kind,
                child_off=child_off, child_idx=u(child_idx), entry_bucket=u(eb), entry_hash=u(eh), entry_set=u(es),
                set_off=u(set_off), set_node_ids=u(set_nodes))


def outcome(header, seq, tree, knobs):
    try:
        p = O.place_sequence(header, seq, tree, knobs["max_iterations"], knobs["min_match_coverage"],
                             knobs["remove_intersection"])
    except O.PlacementError as e:
        return {"error": str(e)}
    return {"status": p.status, "message": p.message, "clade": p.clade, "one": p.one, "rest": p.rest,
            "n_query_kmers": p.n_query_kmers, "n_matched": p.n_matched, "n_root_matched": p.n_root_matched,
            "iterations": p.iterations,
            "response_sha1": hashlib.sha1(json.dumps(O.placement_response(header, p, tree), sort_keys=True)
                                          .encode()).hexdigest()}


def main():
    nwk_name = "Colletotrichum_acutatum_gapdh-PhyML.nwk"
    newick = open(f"{COL}/inputs/{nwk_name}").read()
    tips = O.read_fasta_text(open(f"{COL}/inputs/Colletotrichum_acutatum_gapdh_gapsfree.fasta").read())
    tree = O.tree_from_newick(newick, nwk_name, 70.0)
    O.map_kmers_to_tree(tree, tips, 35, 4)

    # ---- queries ---------------------------------------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(20261018))
    queries = list(tips)
    queries.append(TEST_QUERY)
    gyrb = O.read_fasta_text(open(f"{FD7}/input/bsub-refseq-sample50percentRemaining-clean-diamond.fasta").read())
    queries += gyrb
    for n in range(160):  # sub-reads of the tips, both strands, with substitutions
        name, s = tips[int(rng.integers(len(tips)))]
        ln = int(rng.integers(30, len(s) + 1))
        st = int(rng.integers(0, len(s) - ln + 1))
        sub = list(s[st:st + ln])
        for i in range(len(sub)):
            if rng.random() < 0.02:
                sub[i] = "ACGT"[int(rng.integers(4))]
        sub = "".join(sub)
        if rng.random() < 0.5:
            sub = O.KmersMap.reverse_complement(sub)
        if rng.random() < 0.2:
            sub = sub.lower()
        queries.append((f"sub{n:03d}|{name}|{st}|{ln}", sub))
    for n in range(12):  # chimeras of two tips: stress the one-vs-rest test
        a, b = tips[int(rng.integers(len(tips)))], tips[int(rng.integers(len(tips)))]
        queries.append((f"chim{n:02d}|{a[0]}|{b[0]}", a[1][:120] + b[1][100:]))
    queries.append(("short34", tips[0][1][:34]))
    queries.append(("exact35", tips[0][1][:35]))
    queries.append(("empty", ""))
    queries.append(('quote"back\\slash', "ACGT" * 20))
    queries.append(("random300", "".join("ACGT"[int(x)] for x in rng.integers(0, 4, 300))))
    with open(f"{HERE}/colletotrichum_queries.fasta", "w") as f:
        for h, s in queries:
            f.write(f">{h}\n{s}\n")

    # ---- model -----------------------------------------------------------------------------------
    flat = flatten(tree)
    tree_obj = tree.to_obj()
    tree_obj["kmersMap"] = None
    np.savez_compressed(f"{HERE}/colletotrichum_model.npz", tree_json=np.frombuffer(
        json.dumps(tree_obj).encode(), np.uint8), **flat)

    # ---- expected outcomes -------------------------------------------------------------------------
    exp = {"knobs": KNOBS, "queries": [h for h, _ in queries], "outcomes": {}}
    for kn in KNOBS:
        exp["outcomes"][kn["name"]] = [outcome(h, s, tree, kn) for h, s in queries]
    # full response objects (the record the reference serialises) for one query of each kind
    full, seen_kind = {}, set()
    for (h, s), o in zip(queries, exp["outcomes"]["default"]):
        kind = o.get("status", "Err") + "|" + str(o.get("message"))[:30]
        if "error" in o or (kind in seen_kind and len(full) >= 8):
            continue
        seen_kind.add(kind)
        p = O.place_sequence(h, s, tree)
        full[h] = O.placement_response(h, p, tree)
    exp["responses_default"] = full
    json.dump(exp, open(f"{HERE}/colletotrichum_expected.json", "w"), indent=0, sort_keys=True)

    # ---- reference-pinned facts ----------------------------------------------------------------
    stale = open(f"{COL}/outputs/Colletotrichum_acutatum_gapdh-PhyML.yaml").read()
    head = stale[: stale.index("kmers")] if "kmers" in stale else stale
    leaf_ids = re.findall(r"- id: (\d+)\n\s+name: (\S+)\n\s+kind: LEAF", head)
    res = list(yaml.load_all(open(f"{FD7}/output/result.yaml").read().replace("!", ""), Loader=yaml.BaseLoader))
    first = res[0]
    q0 = dict(gyrb)[first["query"]]
    pins = {
        "murmur3_h1": {"CCAA": 10631256518523097406, "ATAC": 10517626403121597142, "": 0},  # docs/book/02-build-db.md:181,192
        "tree_id": stale.split("\n")[0].split(": ")[1],                                      # outputs/...yaml:1
        "tree_name": nwk_name,
        "stale_golden_leaf_ids": [[int(i), n] for i, n in leaf_ids],                          # pre-order numbering
        "gyrb_first_query": {"header": first["query"], "length": len(q0),
                             "one": int(first["placement"]["one"])},                          # one: 3754 = 2*(1911-34)
        "result_record_keys": {"identity": list(first.keys()),
                               "identity_placement": list(first["placement"].keys()),
                               "identity_clade": [k for k in first["placement"]["clade"].keys()],
                               "codes": sorted({r["code"] for r in res}),
                               "max_resolution_placement_is_scalar": all(
                                   not isinstance(r["placement"], dict) for r in res if r["code"].startswith("Max"))},
    }
    json.dump(pins, open(f"{HERE}/reference_pins.json", "w"), indent=1, sort_keys=True)
    st = {}
    for o in exp["outcomes"]["default"]:
        key = o.get("status", "Err") + (": " + (o.get("message") or "")[:24] if o.get("status") == "Unclassifiable" else "")
        st[key] = st.get(key, 0) + 1
    print(len(queries), "queries;", len(flat["entry_hash"]), "entries;", len(flat["set_off"]) - 1, "sets;", st)


if __name__ == "__main__":
    main()
