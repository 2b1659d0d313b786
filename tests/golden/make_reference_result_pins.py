#!/usr/bin/env python
"""Writes tests/golden/reference_gyrb_results_compact.json.gz from the reference checkout (build container only).

Source: the two result files the reference itself wrote for its bsub-gyrB model
(``tests/data/public/019051d9-...fd7/output/result.yaml``, 14 queries, v0.9.0 layout ``one`` / ``rest``; and
``...fd8/output/result.yaml``, 1 097 queries, the older layout ``oneLen`` / ``restLen``) together with the length of
every query after the reader's filter (``input/*.fasta``).  The k-mer map of that model is not in the reference tree
(Git LFS), so these placements cannot be recomputed - but every record still has to satisfy the rules the descent is
built from, and tests/test_oracle_kats.py holds the oracle's reading of those rules against all of them:
termination (IdentityFound <=> the winner has no non-leaf child; MaxResolutionReached at a node that has some),
the proposal rule ``one > rest``, and ``one <= 2 * (L - k + 1)`` with equality whenever every window of both strands hit.
Compact form: ``{"k": 35, "records": [[file, query length, code, node id, one, rest], ...]}``.
"""
import glob
import gzip
import json
import os
import re
import sys

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import classeq_oracle as O  # noqa: E402

REF = "/root/reference/tests/data/public/019051d9-4c7a-7b2d-9dd1-66ef92236"


def main():
    out = []
    for tag in ("fd7", "fd8"):
        d = REF + tag
        recs = dict(O.read_fasta_text(open(glob.glob(d + "/input/*.fasta")[0]).read()))
        text = re.sub(r"!\w+ ", "", open(d + "/output/result.yaml").read())
        for r in yaml.safe_load_all(text):
            if not r:
                continue
            code, pl = r["code"], r.get("placement")
            node = one = rest = None
            if code == "IdentityFound":
                node = pl["clade"]["id"]
                one, rest = (pl["one"], pl["rest"]) if "one" in pl else (pl["oneLen"], pl["restLen"])
            elif code.startswith("MaxResolutionReached"):
                node = pl
            out.append([tag, len(recs[r["query"]]), code, node, one, rest])
    raw = json.dumps({"k": 35, "records": out}, separators=(",", ":")).encode()
    with open(os.path.join(HERE, "reference_gyrb_results_compact.json.gz"), "wb") as f:
        f.write(gzip.compress(raw, mtime=0))
    print(len(out), "records,", len(raw), "bytes of JSON")


if __name__ == "__main__":
    main()
