"""Comparison helpers shared by the parity tests."""
from classeq2_b200 import _lib

UNCL_BY_MESSAGE_PREFIX = {
    "Query sequence SequenceHeader(": _lib.STATUS_UNCL_NO_MATCH,
    "Query sequence has no overlapping kmers": _lib.STATUS_UNCL_NO_ROOT,
    "Insufficient kmers coverage": _lib.STATUS_UNCL_COVERAGE,
    "Tree introspection not possible": _lib.STATUS_UNCL_NO_INTROSPECTION,
}
ERR_BY_MESSAGE = {
    "The sequence does not contain enough kmers.": _lib.STATUS_ERR_TOO_SHORT,
    "The maximum number of iterations has been reached.": _lib.STATUS_ERR_MAX_ITERATIONS,
    "The root node does not have children. This is unexpected.": _lib.STATUS_ERR_ROOT_NO_CHILDREN,
}


def expected_row(e: dict) -> dict:
    """Oracle outcome (dict as stored in colletotrichum_expected.json / made by outcome_of) ->
    the exact values every cls_result array must hold for that query."""
    if "error" in e:
        st = ERR_BY_MESSAGE[e["error"]]
        row = {"status": st, "node_id": 0, "one": 0, "rest": 0}
        if st == _lib.STATUS_ERR_TOO_SHORT:
            row.update(n_query_kmers=0, n_matched=0, n_root_matched=0, iterations=0)
        return row
    row = {"n_query_kmers": e["n_query_kmers"], "n_matched": e["n_matched"],
           "n_root_matched": e["n_root_matched"], "iterations": e["iterations"], "node_id": 0, "one": 0, "rest": 0}
    if e["status"] == "Unclassifiable":
        for prefix, st in UNCL_BY_MESSAGE_PREFIX.items():
            if e["message"].startswith(prefix):
                row["status"] = st
        if row.get("status") == _lib.STATUS_UNCL_COVERAGE:
            assert e["message"] == f"Insufficient kmers coverage: {e['n_root_matched']}"
    elif e["status"] == "IdentityFound":
        row.update(status=_lib.STATUS_IDENTITY_FOUND, node_id=e["clade"], one=e["one"], rest=e["rest"])
    elif e["status"] == "MaxResolutionReached":
        row.update(status=_lib.STATUS_MAX_RESOLUTION, node_id=e["clade"])
    elif e["status"] == "Inconclusive":
        row.update(status=_lib.STATUS_INCONCLUSIVE)
        row.pop("node_id")
    return row


def outcome_of(oracle, header, seq, tree, max_iterations=None, min_match_coverage=None, remove_intersection=None):
    try:
        p = oracle.place_sequence(header, seq, tree, max_iterations, min_match_coverage, remove_intersection)
    except oracle.PlacementError as err:
        return {"error": str(err)}
    return {"status": p.status, "message": p.message, "clade": p.clade, "one": p.one, "rest": p.rest,
            "n_query_kmers": p.n_query_kmers, "n_matched": p.n_matched, "n_root_matched": p.n_root_matched,
            "iterations": p.iterations}


def assert_rows_equal(res, expected_outcomes, headers=None):
    bad = []
    for i, e in enumerate(expected_outcomes):
        want, got = expected_row(e), res.row(i)
        diff = {k: (got[k], v) for k, v in want.items() if got[k] != v}
        if diff:
            bad.append((i, headers[i] if headers else None, diff))
    assert not bad, f"{len(bad)} of {len(expected_outcomes)} queries differ (got, want): {bad[:5]}"
