"""Synthetic Multi-ATGCN workloads of the shapes BASELINE.json names.

The real DC / Baltimore datasets are not redistributable, so
tests and ``bench.py`` run on seeded synthetic graphs and windows with the same
tensor shapes the LibCity ``MTHDataset`` hands to the model
(``batch['X']`` [B, (len_c+len_p+len_t)*24, N, F], ``batch['y']`` [B, T_out, N, F],
``data_feature`` keys of ``mth_dataset.py:get_data_feature``).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch


class StandardScaler:
    """Same contract as ``libcity/utils/normalization.py:62-76``."""

    def __init__(self, mean, std):
        self.mean = mean
        self.std = std

    def transform(self, data):
        return (data - self.mean) / self.std

    def inverse_transform(self, data):
        return (data * self.std) + self.mean


# name -> (N, batch, T_out, adjtype, adpadj, embed_dim, cheb_order)
WORKLOADS = {
    "dc_multi": dict(N=237, B=64, T_out=3, adjtype="multi", adpadj="bidirection", D=20, cheb=2),
    "dc_adaptive_only": dict(N=237, B=64, T_out=3, adjtype="od", adpadj="bidirection", D=10, cheb=2),
    "baltimore_multi": dict(N=403, B=64, T_out=24, adjtype="multi", adpadj="bidirection", D=20, cheb=2),
    "pems07_scale": dict(N=883, B=256, T_out=12, adjtype="multi", adpadj="bidirection", D=20, cheb=2),
    "tract_8192": dict(N=8192, B=512, T_out=24, adjtype="multi", adpadj="bidirection", D=20, cheb=2),
    "tiny": dict(N=13, B=3, T_out=6, adjtype="multi", adpadj="bidirection", D=4, cheb=2),
}


def make_config(adjtype="multi", adpadj="bidirection", embed_dim=20, cheb_order=2, output_window=24,
                rnn_units=64, num_layers=2, batch_size=64, device="cpu", **extra) -> dict:
    """Config dict with the keys MA.py reads, shipped values
    from ``MultiATGCN.json`` + ``config_user.json`` unless overridden."""
    cfg = dict(input_window=24, output_window=output_window, add_time_in_day=True, add_day_in_week=False,
               node_specific_off=False, fnn_off=False, gcn_off=False, batch_size=batch_size,
               device=torch.device(device) if isinstance(device, str) else device,
               embed_dim_node=embed_dim, embed_dim_adj=embed_dim, adpadj=adpadj, adjtype=adjtype,
               start_dim=0, end_dim=1, load_dynamic=False, rnn_units=rnn_units, num_layers=num_layers,
               cheb_order=cheb_order)
    cfg.update(extra)
    return cfg


def make_data_feature(num_nodes: int, seed: int = 0, len_closeness=2, len_period=1, len_trend=1,
                      input_window=24) -> dict:
    """adj_mx = U(0,1)+0.01 with a dominant positive diagonal (as OD counts have),
    coordinates uniform in a 0.2 x 0.2 degree box around Washington DC, static=None."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    adj = (rng.random((num_nodes, num_nodes)) + 0.01).astype(np.float32)
    adj[np.arange(num_nodes), np.arange(num_nodes)] = adj.max(axis=1) + 1.0
    lon = -77.0 + 0.2 * rng.random(num_nodes)
    lat = 38.8 + 0.2 * rng.random(num_nodes)
    coord = pd.DataFrame({"geo_id": np.arange(num_nodes),
                          "coordinates": ["[%.8f, %.8f]" % (a, b) for a, b in zip(lon, lat)]})
    return {"scaler": StandardScaler(0.0, 1.0), "adj_mx": adj, "static": None, "coordinate": coord,
            "num_nodes": num_nodes, "feature_dim": 2, "output_dim": 1, "ext_dim": 1,
            "len_closeness": len_closeness * input_window, "len_period": len_period * input_window,
            "len_trend": len_trend * input_window, "num_batches": 1}


def make_batch(num_nodes: int, batch: int, output_window: int, seed: int = 0, windows: int = 4,
               device="cpu", pin=False) -> Dict[str, torch.Tensor]:
    """X ~ N(0,1) with channel 1 = time-of-day in [0,1); y ~ N(0,1) (no exact zeros => mask == 1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, windows * 24, num_nodes, 2, generator=g)
    hours = (torch.arange(windows * 24) % 24).float() / 24.0
    x[..., 1] = hours[None, :, None]
    y = torch.randn(batch, output_window, num_nodes, 2, generator=g)
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    return {"X": x.to(device), "y": y.to(device)}


def workload(name: str, seed: int = 0, batch=None, device="cpu") -> Tuple[dict, dict, dict]:
    """(config, data_feature, batch) for a named workload."""
    w = WORKLOADS[name]
    b = w["B"] if batch is None else batch
    cfg = make_config(adjtype=w["adjtype"], adpadj=w["adpadj"], embed_dim=w["D"], cheb_order=w["cheb"],
                      output_window=w["T_out"], batch_size=b, device=device)
    df = make_data_feature(w["N"], seed=seed)
    return cfg, df, make_batch(w["N"], b, w["T_out"], seed=seed, device=device)


def make_series(num_nodes: int, hours: int, seed: int = 0) -> torch.Tensor:
    """A DC-shaped synthetic hourly series ``[hours, N, 2]`` (channel 0 = visits-like non-negative counts, channel 1 = time of
    day in [0, 1)): node-specific scales (heavy-tailed, as README.md:44-53 reports mean 30 / std 84 for DC), a daily and a
    weekly cycle with node-specific phase, a slow spatial coupling through a random sparse graph, and multiplicative noise.
    The real SafeGraph series is not redistributable, so this stands in for it wherever a trained model
    has to be evaluated."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(hours, dtype=torch.float32)
    scale = torch.exp(torch.randn(num_nodes, generator=g) * 1.0 + 2.5)            # median ~12, long right tail
    phase = torch.rand(num_nodes, generator=g) * 4.0 - 2.0                       # hours
    daily = 1.0 + 0.8 * torch.sin(2 * torch.pi * (t[:, None] - 14.0 - phase[None, :]) / 24.0)
    weekly = 1.0 + 0.25 * torch.sin(2 * torch.pi * t[:, None] / (24.0 * 7) + phase[None, :])
    base = scale[None, :] * daily.clamp_min(0.05) * weekly
    # spatial coupling: every node also follows the lagged mean of a few random neighbours
    nbr = torch.randint(num_nodes, (num_nodes, 4), generator=g)
    coupled = base.clone()
    coupled[1:] = 0.7 * base[1:] + 0.3 * base[:-1][:, nbr].mean(-1) * (scale / scale[nbr].mean(-1))[None, :]
    noise = torch.exp(0.15 * torch.randn(hours, num_nodes, generator=g))
    visits = (coupled * noise).clamp_min(0.0)
    out = torch.empty(hours, num_nodes, 2)
    out[..., 0] = visits
    out[..., 1] = ((t % 24) / 24.0)[:, None]
    return out
