"""kb-scale reads against models with k != 35: the one-CTA-per-read geometry with the generic (any k) window hasher.
The two halves are covered elsewhere (generic hasher: test_random_models / test_random_closed_models with short reads;
one CTA per read: test_long_reads_cta_mode with k = 35); this file covers their combination (green on a B200:
profiles/r1h/znew.log)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,m,general", [(21, 3, False), (21, 3, True), (48, 4, False), (12, 0, False)])
def test_long_reads_cta_mode_other_k(k, m, general):
    import classeq2_b200 as cq
    from classeq2_b200 import synth
    from oracle import cpp_oracle
    sm = synth.make_model(40, 1200, 9190 + k, k_size=k, m_size=m)
    lens = np.concatenate([synth.skewed_lengths(120, 78) * 1200 // 1550, np.array([k, k + 1, 280, 300, 301, 640, 1200, 1200])])
    lens = np.clip(lens, k, 1200)
    bases, offsets, _ = synth.make_reads(sm.ref_codes, sm.ref_lens, len(lens), lens, 9192 + k)
    md = cpp_oracle.CppModel.from_flat(sm.flat)
    ix = cq.Index(sm.flat.with_general_sets() if general else sm.flat, device=0)
    for kn in (dict(), dict(remove_intersection=True), dict(min_match_coverage=0.0, max_iterations=2)):
        want = md.place_batch(bases, offsets, kn.get("max_iterations"), kn.get("min_match_coverage"), kn.get("remove_intersection"))
        got = ix.place_batch((bases, offsets), cq.PlaceParams(**kn))
        for f in ("status", "node_id", "one", "rest", "n_query_kmers", "n_matched", "n_root_matched"):
            bad = np.flatnonzero(getattr(got, f) != want[f])
            assert bad.size == 0, (f, kn, bad[:5], getattr(got, f)[bad[:5]], want[f][bad[:5]], lens[bad[:5]])
    ix.close(), md.close()
