"""Shared helpers for the test-suite: golden-file loading and error metrics."""
import glob
import json
import os

import numpy as np
import torch

from multistgraph_b200.synthetic import StandardScaler

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    """-> dict(cfg, data_feature, batch, params, grads, forecast, loss, supports)."""
    import pandas as pd

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = json.loads(str(z["config_json"]))
    cfg["device"] = torch.device("cpu")
    lc, lp, lt = (int(v) for v in z["len_windows"])
    coord = pd.DataFrame({"geo_id": z["geo_id"], "coordinates": [str(s) for s in z["coordinates"]]})
    adj = z["adj_mx"]
    static = z["static"] if "static" in z.files else None   # add_static=true cases carry the node features
    df = {"scaler": StandardScaler(0.0, 1.0), "adj_mx": adj, "static": static, "coordinate": coord,
          "num_nodes": adj.shape[0], "feature_dim": 2, "output_dim": 1, "ext_dim": 1,
          "len_closeness": lc, "len_period": lp, "len_trend": lt, "num_batches": 1}
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    sup = [torch.from_numpy(z[k]) for k in sorted(f for f in z.files if f.startswith("support/"))]
    batch = {"X": torch.from_numpy(z["X"]), "y": torch.from_numpy(z["y"])}
    return dict(cfg=cfg, data_feature=df, batch=batch, params=params, grads=grads,
                forecast=torch.from_numpy(z["forecast"]), loss=float(z["loss"]), supports=sup)


def max_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max(|b|): error relative to the reference tensor's scale.  This is the
    'max rel err' the fp32 bound of 1e-4 (BASELINE.json north_star) is stated in; an
    element-wise ratio is meaningless for gradient entries that are ~0."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.abs().max().item()
    if den == 0.0:
        return (a - b).abs().max().item()
    return ((a - b).abs().max() / den).item()


def elem_rel_err(a: torch.Tensor, b: torch.Tensor, floor_frac: float = 0.05) -> float:
    """Element-wise relative error with an absolute floor: max |a-b| / max(|b|, floor_frac * max|b|).
    Stricter than ``max_rel_err`` (small entries are judged against at most 1/floor_frac of their own size)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = torch.clamp(b.abs(), min=floor_frac * b.abs().max().item())
    if den.max().item() == 0.0:
        return (a - b).abs().max().item()
    return ((a - b).abs() / den).max().item()


def clone_batch(batch, device=None):
    return {k: (v.clone() if device is None else v.clone().to(device)) for k, v in batch.items()}


def exact_pca_lowrank(A, q=None, center=True, niter=2, M=None):
    """Deterministic stand-in for ``torch.pca_lowrank`` (a randomised algorithm the reference calls on every forward
    when ``add_static`` is on, MA.py:291, 407): exact SVD of the centred matrix, columns of V sign-normalised so that
    the result does not depend on device, dtype or RNG state.  The golden ``add_static`` case was generated from the
    real reference with this patch active (tests/golden/make_golden.py); ``tests/conftest.py`` installs it for the suite."""
    A = torch.as_tensor(A)
    q = min(6, A.shape[-2], A.shape[-1]) if q is None else q
    Ac = (A - A.mean(dim=-2, keepdim=True)) if center else A
    U, S, Vh = torch.linalg.svd(Ac.double().cpu(), full_matrices=False)
    V = Vh.transpose(-2, -1)[..., :q]
    idx = V.abs().argmax(dim=-2, keepdim=True)
    sign = torch.sign(torch.gather(V, -2, idx))
    sign[sign == 0] = 1.0
    V = V * sign
    U = U[..., :q] * sign
    return U.to(A.dtype).to(A.device), S[..., :q].to(A.dtype).to(A.device), V.to(A.dtype).to(A.device)
