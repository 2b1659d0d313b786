"""The C-ABI library: loads, exports every symbol include/classeq_b200.h declares, its host-side
helpers agree with the oracle, and compute entry points FAIL LOUDLY without a CUDA device
(no CPU fallback).  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, _has_nvidia_node


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "classeq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cls_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from classeq2_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert _lib.lib.cls_abi_version() == 2


def test_no_torch_in_the_library():
    from classeq2_b200 import _lib
    out = os.popen(f"ldd {_lib.LIB_PATH}").read()
    names = [line.split()[0] for line in out.splitlines() if line.strip()]   # library names only: the load addresses are random hex
    assert names and not [n for n in names if "torch" in n or "c10" in n], names


def test_struct_layouts_match_header():
    from classeq2_b200 import _lib
    assert C.sizeof(_lib.ModelView) == 16 + 8 + 4 * 8 + 8 + 3 * 8 + 8 + 2 * 8
    assert C.sizeof(_lib.Batch) == 24 and C.sizeof(_lib.Params) == 16 and C.sizeof(_lib.Result) == 64
    assert C.sizeof(_lib.LevelCount) == 32 and C.sizeof(_lib.FastaRecords) == 32   # cls_level_count, cls_fasta_records
    p = _lib.Params()
    _lib.lib.cls_params_default(C.byref(p))
    assert (p.max_iterations, p.remove_intersection, p.min_match_coverage) == (1000, 0, 0.7)  # place_sequence.rs:64-75


def test_host_murmur_matches_oracle(oracle, pins):
    import classeq2_b200 as cq
    for s, h in pins["murmur3_h1"].items():
        assert cq.host_murmur3_h1(s.encode()) == h
    rng = np.random.default_rng(7)
    for n in list(range(0, 40)) + [63, 64, 65, 150]:
        b = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        for seed in (0, 1, 2**63 + 5):
            assert cq.host_murmur3_h1(b, seed) == oracle.murmurhash3_x64_128(b, seed)[0]


def test_filter_sequence_matches_oracle(oracle):
    import classeq2_b200 as cq
    for line in ["acgtNNNN-ACGT", "", "xyz", "AcGt\tRYKM*acgu", "ACGTÀé", "ẗẚﬅﬆß", "ｔａ"]:
        assert cq.filter_sequence(line) == oracle.remove_non_iupac_from_sequence(line), repr(line)


def test_model_build_matches_oracle(oracle, col_queries, col_tree, col_npz):
    """cls_model_build (host) == the oracle's map_kmers_to_tree on the Colletotrichum fixture."""
    from classeq2_b200.model import BuiltModel, FlatModel
    z = col_npz
    tflat = FlatModel(35, 4, z["node_id"], z["node_kind"], z["child_off"], z["child_idx"])
    name_to_idx = {}
    clades = list(col_tree.root.walk())
    # FlatModel / golden flattening is the same pre-order as Clade.walk()
    assert [c.id for c in clades] == z["node_id"].tolist()
    for i, c in enumerate(clades):
        if c.is_leaf():
            name_to_idx[c.name] = i
    tips = col_queries[:171]
    tip_node = np.array([name_to_idx[h] for h, _ in tips], np.uint64)
    bases = np.frombuffer("".join(s for _, s in tips).encode(), np.uint8)
    offsets = np.zeros(len(tips) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for _, s in tips])
    bm = BuiltModel(tflat, tip_node, bases, offsets)
    a = bm.arrays()
    got = {}
    for b, h, s in zip(a["entry_bucket"].tolist(), a["entry_hash"].tolist(), a["entry_set"].tolist()):
        got.setdefault(b, {})[h] = set(a["set_node_ids"][int(a["set_off"][s]):int(a["set_off"][s + 1])].tolist())
    assert got == col_tree.kmers_map.map
    assert len(a["set_off"]) - 1 == 226
    bm.close()


@pytest.mark.skipif(_has_nvidia_node(), reason="only meaningful on a machine without a GPU")
def test_compute_fails_loudly_without_gpu(col_flat):
    import classeq2_b200 as cq
    from classeq2_b200 import _lib
    with pytest.raises(_lib.ClsError) as ei:
        cq.Index(col_flat, device=0)
    assert ei.value.code == _lib.CLS_ERR_CUDA
    with pytest.raises(_lib.ClsError):
        cq.debug_kmer_hashes("ACGT" * 20, 35)
    # the sharded-index and peer-buffer entry points as well: no device, no answer
    with pytest.raises(_lib.ClsError) as ei:
        cq.Index(col_flat, device=0, shard=1, n_shards=2)
    assert ei.value.code == _lib.CLS_ERR_CUDA
    # the device model builder too (host-side argument checks come first, then "no device")
    from classeq2_b200.model import BuiltModel, FlatModel
    tflat = FlatModel(35, 4, col_flat.node_id, col_flat.node_kind, col_flat.child_off, col_flat.child_idx)
    seq = np.frombuffer(b"ACGT" * 20, np.uint8).copy()
    with pytest.raises(_lib.ClsError) as ei:
        BuiltModel(tflat, np.array([1], np.uint64), seq, np.array([0, 80], np.uint64), device=0)
    assert ei.value.code == _lib.CLS_ERR_CUDA
    with pytest.raises(_lib.ClsError) as ei:
        BuiltModel(tflat, np.array([1 << 40], np.uint64), seq, np.array([0, 80], np.uint64), device=0)
    assert ei.value.code == _lib.CLS_ERR_INVALID_ARGUMENT
    ptr, h = C.c_void_p(), (C.c_uint8 * 64)()
    assert _lib.lib.cls_peer_alloc(0, 1 << 20, C.byref(ptr), h) == _lib.CLS_ERR_CUDA
    assert _lib.lib.cls_peer_open(0, h, C.byref(ptr)) == _lib.CLS_ERR_CUDA


def test_shard_arguments_are_checked_on_the_host(col_flat):
    import classeq2_b200 as cq
    from classeq2_b200 import _lib
    for shard, n in ((2, 2), (0, 0), (0, 9)):
        with pytest.raises(_lib.ClsError) as ei:
            cq.Index(col_flat, device=0, shard=shard, n_shards=n)
        assert ei.value.code == _lib.CLS_ERR_INVALID_ARGUMENT


def test_index_create_rejects_unsupported_models(col_npz):
    """Rejections happen on the host before any CUDA call, so they are testable without a GPU."""
    import classeq2_b200 as cq
    from classeq2_b200 import _lib
    from classeq2_b200.model import FlatModel
    z = col_npz
    args = [z["node_id"], z["node_kind"], z["child_off"], z["child_idx"]]
    ent = lambda: [z["entry_bucket"].copy(), z["entry_hash"].copy(), z["entry_set"], z["set_off"], z["set_node_ids"]]  # noqa: E731
    with pytest.raises(_lib.ClsError) as ei:          # k = 0
        cq.Index(FlatModel(0, 4, *args, *ent()))
    assert ei.value.code == _lib.CLS_ERR_UNSUPPORTED
    with pytest.raises(_lib.ClsError) as ei:          # m > 12
        cq.Index(FlatModel(35, 13, *args, *ent()))
    assert ei.value.code == _lib.CLS_ERR_UNSUPPORTED
    e = ent()
    e[1][1] = e[1][0]                                 # same hash twice (same or different bucket)
    with pytest.raises(_lib.ClsError) as ei:
        cq.Index(FlatModel(35, 4, *args, *e))
    assert ei.value.code in (_lib.CLS_ERR_UNSUPPORTED, _lib.CLS_ERR_INVALID_ARGUMENT)
    ids = z["node_id"].copy()
    ids[5] = ids[6]                                   # duplicated Clade ids
    with pytest.raises(_lib.ClsError) as ei:
        cq.Index(FlatModel(35, 4, ids, *args[1:], *ent()))
    assert ei.value.code == _lib.CLS_ERR_UNSUPPORTED


def test_host_packer_variants_agree_with_numpy():
    """cls_debug_pack_read: the run-time dispatched body and every named one this CPU can run (portable SWAR,
    AVX2+BMI2, AVX-512) equal a numpy restatement of the layout, flag non-ACGT bytes at every position class, and
    write exactly ceil(len / 16) words (the words after a read belong to the next read, packed by another thread)."""
    from classeq2_b200 import _lib
    rng = np.random.default_rng(5)
    ran = set()
    for L in list(range(0, 200)) + [255, 256, 257, 1000, 1911]:
        codes = rng.integers(0, 4, L)
        s = np.frombuffer(b"ACTG", np.uint8)[codes].copy()          # code order A=0 C=1 T=2 G=3
        lower = rng.random(L) < 0.3
        s[lower] |= 0x20
        nw = (L + 15) // 16
        want = np.full(nw + 2, 0xDEADBEEF, np.uint32)               # sentinels after the read's words
        want[:nw] = 0
        for j, c in enumerate(codes):
            want[j // 16] |= np.uint32(int(c) << (2 * (j % 16)))
        for variant in (0, 1, 2, 3):
            out = np.full(nw + 2, 0xDEADBEEF, np.uint32)
            buf = s if L else np.zeros(1, np.uint8)
            rc = _lib.lib.cls_debug_pack_read(buf.ctypes.data_as(_lib.u8p), L, out.ctypes.data_as(_lib.u32p), len(out), variant)
            if rc == _lib.CLS_ERR_UNSUPPORTED:
                assert variant in (2, 3)
                continue
            ran.add(variant)
            assert rc == 1 and (out == want).all(), (L, variant)
            for pos in ({0, L // 2, L - 1, int(rng.integers(L))} if L else ()):
                for ch in (ord("N"), ord("U"), 0, 0x80 | ord("A"), ord("-")):
                    bad = s.copy()
                    bad[pos] = ch
                    rc = _lib.lib.cls_debug_pack_read(bad.ctypes.data_as(_lib.u8p), L, out.ctypes.data_as(_lib.u32p), len(out), variant)
                    assert rc == 0, (L, variant, pos, ch)
    assert {0, 1} <= ran
    out = np.zeros(4, np.uint32)
    assert _lib.lib.cls_debug_pack_read(out.ctypes.data_as(_lib.u8p), 4, out.ctypes.data_as(_lib.u32p), 4, 7) == _lib.CLS_ERR_INVALID_ARGUMENT

def test_header_is_plain_c99_and_a_c_program_links(tmp_path):
    """include/classeq_b200.h is what a foreign-language binding reads: it must be valid C (not only C++), and a C
    program must link against the library and call into it."""
    import shutil
    import subprocess
    from classeq2_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use.c"
    src.write_text('#include "classeq_b200.h"\n#include <string.h>\n'
                   "int main(void) {\n"
                   "    cls_params p; cls_result r; cls_record_tree t; cls_index *ix = 0;\n"
                   "    memset(&r, 0, sizeof r); memset(&t, 0, sizeof t);\n"
                   "    cls_params_default(&p);\n"
                   "    if (cls_abi_version() != CLS_ABI_VERSION || p.max_iterations != 1000) return 1;\n"
                   "    if (cls_place_batch(ix, 0, &p, &r) != CLS_ERR_INVALID_ARGUMENT) return 2;   /* NULL handle: refused */\n"
                   "    if (!cls_last_error()[0]) return 3;\n"
                   "    return 0;\n}\n")
    exe = tmp_path / "use"
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                        "-L", lib_dir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + lib_dir, "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert subprocess.run([str(exe)]).returncode == 0


def _build_c_example(tmp_path):
    import shutil
    import subprocess
    from classeq2_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe, lib_dir = tmp_path / "place_fasta", os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                        os.path.join(root, "examples", "place_fasta.c"), "-L", lib_dir, "-l:" + os.path.basename(_lib.LIB_PATH),
                        "-Wl,-rpath," + lib_dir, "-lm", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    fa = tmp_path / "q.fasta"
    fa.write_text(">like_a\nACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTCACTGGCCGTCGTTTTACA\n"
                  ">like_d\nACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCGGTTAACCGGTTAACCGTACGTACGATCGATCGGCTAGGT\n>short\nACGT\n")
    return exe, fa


@pytest.mark.skipif(_has_nvidia_node(), reason="only meaningful on a machine without a GPU")
def test_c_example_builds_its_model_and_fails_loudly_without_gpu(tmp_path):
    """examples/place_fasta.c: plain C against the header - model built on the host, then no device, no answer."""
    import subprocess
    exe, fa = _build_c_example(tmp_path)
    r = subprocess.run([str(exe), str(fa), str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 3 and "k-mer entries" in r.stdout and "cls_index_create" in r.stderr
    assert not (tmp_path / "out.yaml").exists()
