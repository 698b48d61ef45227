"""BASELINE shapes against the ORACLE (not against this library's own exact engine).

cfg 1 (DC multi, N=237), cfg 2 (DC adaptive-only, K=2, D=10) and cfg 3 (Baltimore, N=403, 24 -> 24) at batch 8, every
engine mode the library ships (exact fp32 FFMA, TF32 tensor cores, bf16 operand twins - the mode bench.py quotes), forecast
and every parameter gradient compared with the float64 CPU restatement of the reference (oracle/matgcn_oracle.py, itself
pinned on the real MA.py by tests/test_oracle_pin.py).  These are the shapes whose tile counts, half-height tiles, bf16-only
slots and L2 warm-up the small oracle cases never reach.

Two error measures (tests/util.py):
* ``max_rel_err``   = max|a-b| / max|b|            (normwise; the bound the modes are stated in)
* ``elem_rel_err``  = max |a-b| / max(|b|, 0.05 max|b|)   (element-wise with an absolute floor; asserted on the forecast)
"""
import pytest
import torch

from multistgraph_b200 import _cabi
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import workload
from oracle.matgcn_oracle import OracleModel
from tests.util import clone_batch, elem_rel_err, max_rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BATCH = 8
# per mode: (normwise bound on every gradient, normwise bound on the forecast, element-wise bound on the forecast)
BOUNDS = {"exact": (1e-4, 1e-4, 1e-4), "tf32": (1e-2, 2e-3, 1e-2), "bf16": (2e-2, 2e-3, 1e-2)}
_ORACLE = {}


def _oracle(wl):
    """float64 oracle forecast / loss / gradients for a workload (computed once per session: ~10 s at N=403)."""
    if wl not in _ORACLE:
        cfg, df, batch = workload(wl, seed=0, batch=BATCH)
        torch.manual_seed(0)
        params = {k: v.detach().clone() for k, v in MultiATGCN(dict(cfg), df).state_dict().items()}
        ora = OracleModel(cfg, df, params, dtype=torch.float64)
        y = ora.forward(clone_batch(batch)).detach()
        loss = ora.calculate_loss(clone_batch(batch))
        loss.backward()
        _ORACLE[wl] = dict(cfg=cfg, df=df, batch=batch, params=params, y=y, loss=float(loss), grads=ora.grads())
    return _ORACLE[wl]


@pytest.mark.parametrize("mode", ["exact", "tf32", "bf16"])
@pytest.mark.parametrize("wl", ["dc_multi", "dc_adaptive_only", "baltimore_multi"])
def test_baseline_shape_matches_oracle(wl, mode):
    o = _oracle(wl)
    cfg = dict(o["cfg"])
    cfg["device"] = torch.device(DEV)
    cfg["matgcn_mode"] = mode
    model = MultiATGCN(cfg, o["df"]).to(DEV).eval()
    model.load_state_dict(o["params"])
    lib = _cabi.lib()
    before = lib.matgcn_tc_launch_count()
    y = model.predict(clone_batch(o["batch"], DEV))
    loss = model.calculate_loss(clone_batch(o["batch"], DEV))
    loss.backward()
    torch.cuda.synchronize()
    used_tc = lib.matgcn_tc_launch_count() - before
    if mode == "exact":
        # the only tensor-core launches of the exact mode are its dense contractions as 3xTF32 (hi / lo operand split, three
        # k-batches accumulated in fp32: csrc/matgcn.cu prop_3xtf32 - two propagations per step, layer and pass, two forward passes
        # and one backward here - and tb_3xtf32 - the chunked dM contractions); everything else is fp32 FFMA
        assert used_tc <= 3 * (2 * 24 * 2) + 2 * 3 * 24, "exact mode sent other contractions to the tensor-core engine (%d launches)" % used_tc
    else:
        assert used_tc > 20, "fast mode did not run on the tensor-core kernels"
    g_tol, y_tol, y_elem_tol = BOUNDS[mode]
    errs = {"forecast": max_rel_err(y, o["y"]), "forecast_elem": elem_rel_err(y, o["y"]),
            "loss": abs(loss.item() - o["loss"]) / abs(o["loss"])}
    for k, p in model.named_parameters():
        ref = o["grads"].get(k)
        if ref is not None:
            assert p.grad is not None, k
            errs["d" + k] = max_rel_err(p.grad, ref)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("[oracle %s %s B=%d] " % (wl, mode, BATCH) + ", ".join("%s=%.2e" % kv for kv in worst))
    assert errs["forecast"] < y_tol, errs["forecast"]
    assert errs["forecast_elem"] < y_elem_tol, errs["forecast_elem"]
    assert errs["loss"] < y_tol
    bad = {k: v for k, v in errs.items() if k.startswith("d") and not (v < g_tol)}
    assert not bad, bad
