"""Pin the CPU oracle: (i) against the frozen vectors generated from the real MA.py,
(ii) against the real MA.py imported live when /root/reference is mounted."""
import os
import sys

import pytest
import torch

from oracle.matgcn_oracle import OracleModel
from tests.util import clone_batch, golden_names, load_golden, max_rel_err

TOL = 2e-5  # fp32 re-association noise only: same algorithm, same op order class


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    g = load_golden(name)
    m = OracleModel(g["cfg"], g["data_feature"], g["params"])
    for got, want in zip(m.static_sup, g["supports"]):
        assert max_rel_err(got, want) < 1e-6
    y = m.forward(clone_batch(g["batch"]))
    assert y.shape == g["forecast"].shape
    assert max_rel_err(y, g["forecast"]) < TOL
    loss = m.calculate_loss(clone_batch(g["batch"]))
    assert abs(loss.item() - g["loss"]) < TOL * max(1.0, abs(g["loss"]))
    loss.backward()
    grads = m.grads()
    assert g["grads"], "golden file holds no gradients"
    for k, want in g["grads"].items():
        assert grads[k] is not None, k
        assert max_rel_err(grads[k], want) < 5e-5, k
    # parameters the reference leaves without a gradient stay without one here too
    for k, v in grads.items():
        if k not in g["grads"] and v is not None:
            assert v.abs().max().item() == 0.0, k


@pytest.mark.skipif(not os.path.isdir("/root/reference/libcity"), reason="reference mount absent (GPU box)")
@pytest.mark.parametrize("adjtype,adpadj,cheb", [("multi", "bidirection", 2), ("od", "bidirection", 2),
                                                  ("multi", "none", 3), ("od", "unidirection", 1)])
def test_oracle_matches_live_reference(adjtype, adpadj, cheb):
    sys.path.insert(0, "/root/reference")
    from libcity.model.traffic_flow_prediction.MultiATGCN import MultiATGCN as RefModel
    from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature

    cfg = make_config(adjtype=adjtype, adpadj=adpadj, embed_dim=6, cheb_order=cheb, output_window=24,
                      rnn_units=16, batch_size=2)
    df = make_data_feature(17, seed=7)
    batch = make_batch(17, 2, 24, seed=7)
    torch.manual_seed(3)
    ref = RefModel(dict(cfg), df).eval()
    loss_ref = ref.calculate_loss(clone_batch(batch))
    loss_ref.backward()
    ora = OracleModel(cfg, df, ref.state_dict())
    loss = ora.calculate_loss(clone_batch(batch))
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < TOL
    got = ora.grads()
    for k, p in ref.named_parameters():
        if p.grad is None:
            continue
        assert max_rel_err(got[k], p.grad) < 5e-5, k
