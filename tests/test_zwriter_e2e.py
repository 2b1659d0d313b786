"""place_sequences end to end with the library's native record writer against the Python emitters: the two runs write
byte-identical result and error files (YAML and JSON lines, with annotations).  The writer itself is covered without a
GPU in tests/test_record_writer.py; this file sorts last among the GPU tests on purpose."""
import os

import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("fmt", ["yaml", "jsonl"])
def test_native_and_python_writers_write_the_same_files(tmp_path, col_tree, fmt):
    import classeq2_b200 as cq
    from classeq2_b200 import placement as ps
    tree = cq.Tree.from_obj(col_tree.to_obj())
    tree.annotations = [{"clade": 0, "meta": [ps.Tag("SciName", "Colletotrichum acutatum complex"), ps.Tag("Taxid", 27357)]},
                        {"clade": 18, "meta": [ps.Tag("Note", "two\nlines\n")]}, {"clade": 1}]
    fa = os.path.join(GOLDEN, "colletotrichum_queries.fasta")
    index = cq.Index(tree, device=0)
    a, b = tmp_path / "n" / "r", tmp_path / "p" / "r"
    ta = cq.place_sequences(fa, tree, a, output_format=fmt, index=index, writer="native")
    tb = cq.place_sequences(fa, tree, b, output_format=fmt, index=index, writer="python")
    index.close()
    assert [t.sequence for t in ta] == [t.sequence for t in tb] and len(ta) > 300
    fa_, fb_ = tmp_path / "n" / f"r.{fmt}", tmp_path / "p" / f"r.{fmt}"
    assert fa_.read_bytes() == fb_.read_bytes() and len(fa_.read_bytes()) > 10000 and b"annotations" in fa_.read_bytes()
    assert (tmp_path / "n" / "r.error").read_bytes() == (tmp_path / "p" / "r.error").read_bytes()


@pytest.mark.parametrize("fmt", ["yaml", "jsonl"])
def test_one_call_place_sequences_writes_the_same_files(tmp_path, col_tree, fmt):
    """cls_place_sequences (reader, cls_place_batch, writer and the file handling of mod.rs:73-106 in one library call)
    against the Python driver; its two host halves are covered without a GPU in tests/test_record_writer.py."""
    import classeq2_b200 as cq
    tree = cq.Tree.from_obj(col_tree.to_obj())
    fa = os.path.join(GOLDEN, "colletotrichum_queries.fasta")
    index = cq.Index(tree, device=0)
    n = cq.place_sequences_native(fa, tree, tmp_path / "c" / "r.out", output_format=fmt, index=index, remove_intersection=True)
    t = cq.place_sequences(fa, tree, tmp_path / "p" / "r.out", output_format=fmt, index=index, remove_intersection=True, writer="python")
    assert n == len(t) > 300
    assert (tmp_path / "c" / f"r.{fmt}").read_bytes() == (tmp_path / "p" / f"r.{fmt}").read_bytes()
    assert (tmp_path / "c" / "r.error").read_bytes() == (tmp_path / "p" / "r.error").read_bytes()
    with pytest.raises(FileExistsError):
        cq.place_sequences_native(fa, tree, tmp_path / "c" / "r.out", output_format=fmt, index=index)
    assert cq.place_sequences_native(fa, tree, tmp_path / "c" / "r.out", output_format=fmt, index=index, overwrite=True) == n
    index.close()


def test_c_example_places_a_fasta_file(tmp_path):
    """examples/place_fasta.c (plain C: cls_model_build, cls_index_create, cls_place_sequences) on a B200; the expected
    placements are the oracle's for the same four-tip model."""
    import subprocess
    import yaml
    from test_abi import _build_c_example
    exe, fa = _build_c_example(tmp_path)
    r = subprocess.run([str(exe), str(fa), str(tmp_path / "out.any")], capture_output=True, text=True)
    assert r.returncode == 0 and "3 queries placed" in r.stdout, (r.stdout, r.stderr)
    recs = list(yaml.safe_load_all((tmp_path / "out.yaml").read_text()))
    assert [x["query"] for x in recs] == ["like_a", "like_d"] and all(x["code"] == "IdentityFound" for x in recs)
    assert recs[0]["placement"]["clade"]["id"] == 1 and recs[1]["placement"]["clade"]["id"] == 4
    assert (recs[0]["placement"]["one"], recs[0]["placement"]["rest"]) == (96, 16)
    assert [c["name"] for c in recs[0]["placement"]["clade"]["children"]] == ["tip_a", "tip_b"]
    assert (tmp_path / "out.error").read_text() == "The sequence does not contain enough kmers."
