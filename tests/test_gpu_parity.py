"""Parity of the CUDA path (through the C ABI) with the oracle / golden vectors.  Needs a GPU.

fp32 mode bound (BASELINE.json north_star): max rel err 1e-4 on forecasts and gradients,
where rel err = max|a-b| / max|b| per tensor (tests/util.py).
"""
import ctypes

import numpy as np
import pytest
import torch

from multistgraph_b200 import _cabi, ops
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature
from oracle.matgcn_oracle import OracleModel
from tests import host_mirror as hm
from tests.util import clone_batch, golden_names, load_golden, max_rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
ACCEL = golden_names()
DEV = "cuda:0"
FLAGS = 0  # exact mode


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def _report(tag, errs):
    bad = {k: v for k, v in errs.items() if not (v < TOL)}
    print("[%s] " % tag + ", ".join("%s=%.2e" % kv for kv in errs.items()))
    assert not bad, "%s: buffers over tolerance: %s" % (tag, bad)


# ------------------------------------------------------------------------------------------
# operator level: every C-ABI entry point against the torch mirror, buffer by buffer
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,D", [(5, 3), (37, 10), (403, 20), (1000, 7)])
def test_adaptive_adjacency_op(N, D):
    L, Rt = _rand(N, D, seed=1), _rand(N, D, seed=2)
    ldm = (N + 7) // 8 * 8
    Lc, Rc = L.clone().requires_grad_(True), Rt.clone().requires_grad_(True)
    A_ref = hm.MirrorAdjFn.apply(Lc, Rc, ldm)
    dA = _rand(N, ldm, seed=3)
    A_ref.backward(dA)
    Lg, Rg = L.to(DEV).requires_grad_(True), Rt.to(DEV).requires_grad_(True)
    A = ops.adaptive_adjacency(Lg, Rg, ldm)
    A.backward(dA.to(DEV))
    _report("adj N=%d" % N, {"A": max_rel_err(A, A_ref), "dL": max_rel_err(Lg.grad, Lc.grad),
                             "dRt": max_rel_err(Rg.grad, Rc.grad)})
    assert A[:, N:].abs().max().item() == 0.0 if ldm > N else True


@pytest.mark.parametrize("N,D,K,I,O", [(4, 2, 2, 3, 5), (33, 10, 5, 66, 128), (50, 20, 5, 128, 64), (7, 1, 3, 10, 16)])
def test_node_weights_op(N, D, K, I, O):
    E, pool, bp = _rand(N, D, seed=1), _rand(D, K, I, O, seed=2, scale=0.1), _rand(D, O, seed=3)
    c = torch.softmax(_rand(K, seed=4), 0)
    dW, db = _rand(N, K, I, O, seed=5), _rand(N, O, seed=6)
    ref_in = [t.clone().requires_grad_(True) for t in (E, pool, bp, c)]
    W_ref, b_ref = hm.MirrorNodeWeightsFn.apply(*ref_in)
    torch.autograd.backward([W_ref, b_ref], [dW, db])
    gpu_in = [t.to(DEV).requires_grad_(True) for t in (E, pool, bp, c)]
    W, b = ops.node_weights(*gpu_in)
    torch.autograd.backward([W, b], [dW.to(DEV), db.to(DEV)])
    errs = {"W": max_rel_err(W, W_ref), "b": max_rel_err(b, b_ref)}
    for nm, a, r in zip(["dE", "dpool", "dbias_pool", "dc"], gpu_in, ref_in):
        errs[nm] = max_rel_err(a.grad, r.grad)
    _report("nodeweights N=%d" % N, errs)


LAYER_SHAPES = [
    # T, N, B, Cin, H, Kp, n_adp, h0
    (3, 5, 2, 2, 8, 1, 1, False),
    (4, 19, 3, 2, 16, 4, 1, True),
    (24, 37, 5, 64, 64, 4, 1, False),
    (5, 130, 70, 3, 32, 2, 0, False),
    (2, 150, 9, 64, 64, 8, 2, True),
]


@pytest.mark.parametrize("T,N,B,Cin,H,Kp,n_adp,with_h0", LAYER_SHAPES)
def test_encoder_layer_op_buffers(T, N, B, Cin, H, Kp, n_adp, with_h0):
    K, I = Kp + 1, Cin + H
    ldm = (N + 7) // 8 * 8
    x = _rand(T, N, B, Cin, seed=1)
    h0 = _rand(N, B, H, seed=2, scale=0.5) if with_h0 else None
    M = torch.zeros(Kp, N, ldm)
    M[:, :, :N] = torch.softmax(_rand(Kp, N, N, seed=3), dim=2) * torch.tensor([1.0, -1.0] * Kp)[:Kp].view(Kp, 1, 1)
    s = 1.0 / np.sqrt(K * I)
    Wg, Wu = _rand(N, K, I, 2 * H, seed=4, scale=s), _rand(N, K, I, H, seed=5, scale=s)
    bg, bu = _rand(N, 2 * H, seed=6, scale=0.1), _rand(N, H, seed=7, scale=0.1)
    r = 1.0 / np.sqrt(I)
    Rgw, Ruw = _rand(2 * H, I, seed=8, scale=r), _rand(H, I, seed=9, scale=r)
    Rgb, Rub = _rand(2 * H, seed=10, scale=0.1), _rand(H, seed=11, scale=0.1)
    mix = torch.sigmoid(_rand(T, seed=12))
    dY = _rand(T, N, B, H, seed=13)

    y_ref, sv = hm.layer_fwd(x, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix)
    g_ref = hm.layer_bwd(dY, sv, M, Wg, Wu, Rgw, Ruw, mix, n_adp, with_h0)

    L = _cabi.lib()
    dims = (T, N, B, Cin, H, K)
    d = lambda t: None if t is None else t.to(DEV).contiguous()  # noqa: E731
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    xd, h0d, Md, Wgd, bgd, Wud, bud = d(x), d(h0), d(M), d(Wg), d(bg), d(Wu), d(bu)
    Rgwd, Rgbd, Ruwd, Rubd, mixd, dYd = d(Rgw), d(Rgb), d(Ruw), d(Rub), d(mix), d(dY)
    ws = torch.zeros(L.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(L.matgcn_encoder_layer_fwd(*dims, ldm, p(xd), xd.stride(0), p(h0d), p(Md), p(Wgd), p(bgd), p(Wud),
                                           p(bud), p(Rgwd), p(Rgbd), p(Ruwd), p(Rubd), p(mixd), p(ws), FLAGS, st), "fwd")
    torch.cuda.synchronize()

    def slot(name, shape):
        off = L.matgcn_encoder_layer_slot_offset(name.encode(), *dims)
        assert off != ctypes.c_size_t(-1).value
        n = int(np.prod(shape))
        return ws[off:off + n].view(*shape).cpu()

    U = (N, B, H)
    errs = {"PX": max_rel_err(slot("PX", (T, K, N, B, Cin)), sv["PX"]),
            "GX": max_rel_err(slot("GX", (T, N, B, 3 * H)), sv["GX"]),
            "RX": max_rel_err(slot("RX", (T, N, B, 3 * H)), sv["RX"]),
            "PZ": max_rel_err(slot("PZ", (T, K) + U), sv["PZ"])}
    ph = slot("PH", (T * K + 1,) + U)
    errs["PH"] = max_rel_err(ph[:T * K].view(T, K, *U), sv["PH"][:T])
    errs["y_last"] = max_rel_err(ph[T * K], sv["PH"][T, 0])
    for nm in ["Z", "R", "HC", "H1", "Z2", "R2", "HC2", "ZH2"]:
        errs[nm] = max_rel_err(slot(nm, (T,) + U), sv[nm])
    _report("layer fwd", errs)

    bws = torch.zeros(L.matgcn_encoder_layer_bwd_ws_bytes(*dims, n_adp) // 4, device=DEV)
    new = lambda *s: torch.full(s, float("nan"), device=DEV)  # noqa: E731
    dx, dM = new(T, N, B, Cin), new(Kp, N, ldm)
    dh0 = new(N, B, H) if with_h0 else None
    dWg, dbg, dWu, dbu = new(N, K, I, 2 * H), new(N, 2 * H), new(N, K, I, H), new(N, H)
    dRgw, dRgb, dRuw, dRub, dmix = new(2 * H, I), new(2 * H), new(H, I), new(H), new(T)
    _cabi.check(L.matgcn_encoder_layer_bwd(*dims, ldm, n_adp, p(dYd), dYd.stride(0), p(Md), p(Wgd), p(Wud), p(Rgwd),
                                           p(Ruwd), p(mixd), p(ws), p(bws), p(dx), p(dh0), p(dM), p(dWg), p(dbg),
                                           p(dWu), p(dbu), p(dRgw), p(dRgb), p(dRuw), p(dRub), p(dmix), FLAGS, st), "bwd")
    torch.cuda.synchronize()
    errs = {"DG": max_rel_err(slot("GX", (T, N, B, 3 * H)), g_ref["DG"]),
            "DR": max_rel_err(slot("RX", (T, N, B, 3 * H)), g_ref["DR"]),
            "dx": max_rel_err(dx, g_ref["dX"]), "dM": max_rel_err(dM, g_ref["dM"]),
            "dWg": max_rel_err(dWg, g_ref["dWg"]), "dbg": max_rel_err(dbg, g_ref["dbg"]),
            "dWu": max_rel_err(dWu, g_ref["dWu"]), "dbu": max_rel_err(dbu, g_ref["dbu"]),
            "dRgw": max_rel_err(dRgw, g_ref["dRgw"]), "dRgb": max_rel_err(dRgb, g_ref["dRgb"]),
            "dRuw": max_rel_err(dRuw, g_ref["dRuw"]), "dRub": max_rel_err(dRub, g_ref["dRub"]),
            "dmix": max_rel_err(dmix, g_ref["dmix"])}
    if with_h0:
        errs["dh0"] = max_rel_err(dh0, g_ref["dh0"])
    _report("layer bwd", errs)


# ------------------------------------------------------------------------------------------
# model level: drop-in MultiATGCN on the GPU against the frozen reference vectors / the oracle
# ------------------------------------------------------------------------------------------
def _gpu_model(cfg, df, params):
    cfg = dict(cfg)
    cfg["device"] = torch.device(DEV)
    model = MultiATGCN(cfg, df).to(DEV).eval()
    model.load_state_dict(params)
    return model


@pytest.mark.parametrize("name", ACCEL)
def test_model_matches_golden(name):
    g = load_golden(name)
    model = _gpu_model(g["cfg"], g["data_feature"], g["params"])
    y = model.predict(clone_batch(g["batch"], DEV))
    loss = model.calculate_loss(clone_batch(g["batch"], DEV))
    loss.backward()
    errs = {"forecast": max_rel_err(y, g["forecast"]), "loss": abs(loss.item() - g["loss"]) / abs(g["loss"])}
    for k, p in model.named_parameters():
        if k in g["grads"]:
            assert p.grad is not None, k
            errs["d" + k] = max_rel_err(p.grad, g["grads"][k])
        else:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, k
    _report("golden " + name, errs)


@pytest.mark.parametrize("N,B,adjtype,adpadj,D,tout", [(45, 6, "multi", "bidirection", 20, 24),
                                                        (60, 4, "od", "bidirection", 10, 3),
                                                        (33, 3, "multi", "none", 20, 12)])
def test_model_matches_oracle_default_width(N, B, adjtype, adpadj, D, tout):
    """rnn_units=64 (the shipped width), seeded weights; oracle in float64 as the yardstick."""
    cfg = make_config(adjtype=adjtype, adpadj=adpadj, embed_dim=D, output_window=tout, batch_size=B)
    df = make_data_feature(N, seed=5)
    batch = make_batch(N, B, tout, seed=5)
    torch.manual_seed(0)
    model = _gpu_model(cfg, df, MultiATGCN(dict(cfg), df).state_dict())
    ora = OracleModel(cfg, df, {k: v.cpu() for k, v in model.state_dict().items()}, dtype=torch.float64)
    y_ref = ora.forward(clone_batch(batch))
    loss_ref = ora.calculate_loss(clone_batch(batch))
    loss_ref.backward()
    y = model.predict(clone_batch(batch, DEV))
    loss = model.calculate_loss(clone_batch(batch, DEV))
    loss.backward()
    errs = {"forecast": max_rel_err(y, y_ref), "loss": abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())}
    grads = ora.grads()
    for k, p in model.named_parameters():
        if grads.get(k) is not None and p.grad is not None:
            errs["d" + k] = max_rel_err(p.grad, grads[k])
    _report("oracle N=%d" % N, errs)


# ------------------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties (the oracle cannot run these in seconds)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wl", ["baltimore_multi"])
def test_full_size_batch_split_property(wl):
    """Samples are independent: the forecast of a batch equals the forecasts of its two halves,
    and the summed parameter gradients of the halves equal the gradient of the whole batch."""
    from multistgraph_b200.synthetic import workload
    cfg, df, batch = workload(wl, seed=0, batch=16)
    torch.manual_seed(0)
    model = _gpu_model(cfg, df, MultiATGCN(dict(cfg), df).state_dict())

    def run(lo, hi):
        model.zero_grad(set_to_none=True)
        sub = {k: v[lo:hi].clone().to(DEV) for k, v in batch.items()}
        y = model.predict(sub)
        tgt = sub["y"][..., :1]
        ((y - tgt) ** 2).sum().backward()
        return y.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    y_all, g_all = run(0, 16)
    y_a, g_a = run(0, 8)
    y_b, g_b = run(8, 16)
    assert torch.isfinite(y_all).all()
    errs = {"forecast": max_rel_err(torch.cat([y_a, y_b], 0), y_all)}
    for k in g_all:
        errs["d" + k] = max_rel_err(g_a[k] + g_b[k], g_all[k])
    _report("split " + wl, errs)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libmatgcn.so")
    with pytest.raises(_cabi.MatgcnError):
        ops.adaptive_adjacency(torch.zeros(4, 2, device=DEV), torch.zeros(4, 2, device=DEV), 8)
