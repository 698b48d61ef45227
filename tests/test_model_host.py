"""Host-side logic of the drop-in model on CPU: parameter contract, graph prep, fusion, head
and loss - with the three C-ABI operators swapped (tests only) for the torch mirror of the
kernel-level algorithm, so that the hand-derived backward is also checked against the golden
gradients of the real reference before any GPU run."""
import os
import sys

import numpy as np
import pytest
import torch

from multistgraph_b200 import graph_prep, ops
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import make_config, make_data_feature
from oracle import matgcn_oracle
from tests import host_mirror
from tests.util import clone_batch, golden_names, load_golden, max_rel_err

ACCEL = golden_names()


@pytest.fixture()
def mirrored_ops():
    restore = host_mirror.install(ops)
    yield
    restore()


@pytest.mark.parametrize("name", ACCEL)
def test_model_with_mirror_matches_golden(name, mirrored_ops):
    g = load_golden(name)
    model = MultiATGCN(dict(g["cfg"]), g["data_feature"]).eval()
    missing, unexpected = model.load_state_dict(g["params"], strict=True)
    assert not missing and not unexpected
    y = model.predict(clone_batch(g["batch"]))
    assert y.shape == g["forecast"].shape
    assert max_rel_err(y, g["forecast"]) < 2e-5
    loss = model.calculate_loss(clone_batch(g["batch"]))
    assert abs(loss.item() - g["loss"]) < 2e-5
    loss.backward()
    for k, p in model.named_parameters():
        if k in g["grads"]:
            assert p.grad is not None, k
            assert max_rel_err(p.grad, g["grads"][k]) < 1e-4, k
        else:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, k


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_contract(name):
    """Same parameter names, shapes and registration order as the reference checkpoint."""
    g = load_golden(name)
    model = MultiATGCN(dict(g["cfg"]), g["data_feature"])
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["params"].keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["params"][k].shape), k
    assert [k for k, _ in model.named_parameters()] == [k for k in g["params"].keys()]


def test_static_supports_match_oracle():
    for adjtype in ["multi", "od", "dist", "cosine", "identity"]:
        df = make_data_feature(23, seed=3)
        cfg = make_config(adjtype=adjtype)
        ora = matgcn_oracle.build_static_supports(cfg, df)
        mine = graph_prep.static_views(adjtype, df)
        assert len(ora["supports"]) == len(mine["laplacians"])
        for a, b in zip(ora["supports"], mine["laplacians"]):
            assert max_rel_err(torch.from_numpy(b), a) < 1e-6
        assert max_rel_err(torch.from_numpy(np.asarray(mine["adj_mx"])), ora["adj_mx"]) < 1e-6


def test_unsorted_geo_ids_follow_pivot_order():
    df = make_data_feature(9, seed=5)
    perm = np.random.default_rng(0).permutation(9)
    df["coordinate"] = df["coordinate"].iloc[perm].reset_index(drop=True)
    a = graph_prep.distance_adjacency(df["coordinate"])
    b = matgcn_oracle.distance_view(df["coordinate"])
    assert np.allclose(a, b, atol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference/libcity"), reason="reference mount absent (GPU box)")
def test_seeded_init_matches_reference():
    """A seeded construction draws the same initial weights as the reference (same RNG order)."""
    sys.path.insert(0, "/root/reference")
    from libcity.model.traffic_flow_prediction.MultiATGCN import MultiATGCN as RefModel

    for extra in [dict(), dict(fnn_off=True), dict(node_specific_off=True), dict(adjtype="od", cheb_order=3)]:
        kw = dict(adjtype="multi", adpadj="bidirection", embed_dim=5, rnn_units=8, output_window=6)
        kw.update(extra)
        cfg = make_config(**kw)
        df = make_data_feature(10, seed=2)
        torch.manual_seed(11)
        ref = RefModel(dict(cfg), df)
        torch.manual_seed(11)
        mine = MultiATGCN(dict(cfg), df)
        rs, ms = ref.state_dict(), mine.state_dict()
        assert list(rs.keys()) == list(ms.keys())
        for k in rs:
            assert torch.equal(rs[k], ms[k]), k


def test_cpu_tensor_raises_without_fallback():
    """No CPU fallback: feeding the real operators a CPU tensor must fail loudly."""
    g = load_golden("od_bi_k2")
    model = MultiATGCN(dict(g["cfg"]), g["data_feature"]).eval()
    model.load_state_dict(g["params"])
    with pytest.raises(Exception):
        model.predict(clone_batch(g["batch"]))


def test_similarity_view_duplicate_and_near_duplicate_rows():
    """MA.py:246-248: ``cdist`` distances, exact zeros -> 1.  Duplicated rows of unnormalised features (magnitudes up to 1e5)
    must hit the ``== 0 -> 1`` rule and rows 1e-3 apart must not lose digits (the Gram expansion did both wrong)."""
    import numpy as np
    from scipy.spatial.distance import cdist

    from multistgraph_b200 import graph_prep

    for seed in range(6):
        rng = np.random.default_rng(seed)
        n = 97
        s = rng.random((n, 30)) * 10.0 ** rng.integers(0, 6, size=(1, 30))
        s[5] = s[40]
        s[17] = s[40]
        s[60] = s[61] + 1e-3
        ref = cdist(s, s, metric="euclidean")
        ref[ref == 0] = 1
        ref = (1.0 / ref).astype(np.float32)
        got = graph_prep.similarity_view(s, n)
        assert got[5, 40] == 1.0 and got[40, 17] == 1.0 and got[5, 17] == 1.0
        assert np.array_equal(got, ref)
