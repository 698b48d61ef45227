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
    df = {"scaler": StandardScaler(0.0, 1.0), "adj_mx": adj, "static": None, "coordinate": coord,
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


def clone_batch(batch, device=None):
    return {k: (v.clone() if device is None else v.clone().to(device)) for k, v in batch.items()}
