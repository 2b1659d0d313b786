"""Shared fixtures.  `-m "not gpu"` must pass on a machine without a GPU; `-m gpu` tests call
the sm_100a kernels through the C ABI and compare them with the CPU oracle (oracle/)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_nvidia_node() -> bool:
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    # On a box with no NVIDIA device node at all (the build container) GPU tests are skipped; on
    # a GPU box they run and fail loudly if the CUDA path is unusable - never a silent fallback.
    if _has_nvidia_node():
        return
    skip = pytest.mark.skip(reason="no NVIDIA device node on this machine")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The in-tree shared objects are build artefacts (git-ignored): `make` them (incremental - a library older than
    its sources is rebuilt, never tested)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import classeq_oracle
    return classeq_oracle


@pytest.fixture(scope="session")
def pins():
    return json.load(open(os.path.join(GOLDEN, "reference_pins.json")))


@pytest.fixture(scope="session")
def col_queries(oracle):
    """(header, body) records of the committed query FASTA, read with the oracle's FASTA reader."""
    text = open(os.path.join(GOLDEN, "colletotrichum_queries.fasta")).read()
    recs = oracle.read_fasta_text(text)
    # the FASTA reader drops the trailing empty record and sends mid-file empty ones: the file
    # holds one empty-body query ("empty") which the reader emits as ("empty", "")
    return recs


@pytest.fixture(scope="session")
def col_expected():
    return json.load(open(os.path.join(GOLDEN, "colletotrichum_expected.json")))


@pytest.fixture(scope="session")
def col_npz():
    z = np.load(os.path.join(GOLDEN, "colletotrichum_model.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def col_tree(oracle, col_npz):
    """The Colletotrichum model as an oracle ``Tree`` (tree JSON + k-mer map from the flat arrays)."""
    obj = json.loads(bytes(col_npz["tree_json"]).decode())
    tree = oracle.Tree.from_obj(obj)
    km = oracle.KmersMap(int(col_npz["k_size"]), int(col_npz["m_size"]))
    so, sn = col_npz["set_off"], col_npz["set_node_ids"]
    for b, h, s in zip(col_npz["entry_bucket"].tolist(), col_npz["entry_hash"].tolist(), col_npz["entry_set"].tolist()):
        km.map.setdefault(b, {})[h] = set(sn[int(so[s]):int(so[s + 1])].tolist())
    tree.kmers_map = km
    return tree


@pytest.fixture(scope="session")
def col_flat(col_npz):
    """The same model as the product's FlatModel (cls_model_view)."""
    from classeq2_b200.model import FlatModel
    z = col_npz
    return FlatModel(int(z["k_size"]), int(z["m_size"]), z["node_id"], z["node_kind"], z["child_off"], z["child_idx"],
                     z["entry_bucket"], z["entry_hash"], z["entry_set"], z["set_off"], z["set_node_ids"])
