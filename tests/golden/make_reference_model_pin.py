#!/usr/bin/env python
"""Writes tests/golden/reference_built_model_k12.json.gz from the reference checkout (build container only).

Source: ``core/src/tests/data/colletotrichum-acutatom-complex/outputs/Colletotrichum_acutatum_gapdh-PhyML.yaml`` -
a model WRITTEN BY THE REFERENCE ITSELF (an early version: ``kSize: 12``, no minimizers, the map keyed by the
k-mer string instead of its hash, forward strand only - the reverse complement came with v0.2.3,
CHANGELOG.md:177-179) from the newick + FASTA inputs next to it.  It is the one artefact in the reference tree
that holds k-mer -> node-set data, so it pins, against the reference's own output:

* the tree that ``Tree::init_from_file`` builds from the newick (ids, kinds, names, supports, lengths, shape),
* the windowing (every forward window of length k),
* the node set of a k-mer (union of the root -> tip id paths, both ends included),
* the header / sequence pairing of the MSA loop (header i is indexed with sequence i-1; the last one is dropped).

The file is re-serialised compactly (same content): ``{"k_size", "root", "kmers": {kmer: sorted node ids}}``.
"""
import gzip
import json
import os

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/core/src/tests/data/colletotrichum-acutatom-complex/outputs/Colletotrichum_acutatum_gapdh-PhyML.yaml"


def norm(node: dict) -> dict:
    out = {}
    for k, v in node.items():
        if k == "children":
            out[k] = [norm(c) for c in v]
        elif k in ("length", "support"):
            out[k] = float(v)          # PyYAML reads "1e-8" (no dot) as a string
        else:
            out[k] = v
    return out


def main():
    import shutil
    # the aligned (MAFFT: lower case, '-' gaps) version of the same 171 sequences, as the reference ships it: what its own
    # build test feeds to the reader's filter (core/src/use_cases/build_database/mod.rs:189-208)
    shutil.copy(os.path.dirname(os.path.dirname(SRC)) + "/inputs/Colletotrichum_acutatum_gapdh_mafft.fasta",
                os.path.join(HERE, "Colletotrichum_acutatum_gapdh_mafft.fasta"))
    d = yaml.safe_load(open(SRC))
    pin = {"source": SRC.replace("/root/reference/", ""), "id": d["id"], "name": d["name"],
           "k_size": int(d["kmersMap"]["kSize"]), "root": norm(d["root"]),
           "kmers": {k: sorted(int(x) for x in v) for k, v in sorted(d["kmersMap"]["map"].items())}}
    raw = json.dumps(pin, separators=(",", ":"), sort_keys=True).encode()
    with open(os.path.join(HERE, "reference_built_model_k12.json.gz"), "wb") as f:
        f.write(gzip.compress(raw, mtime=0))
    print(len(pin["kmers"]), "k-mers,", len(raw), "bytes of JSON")


if __name__ == "__main__":
    main()
