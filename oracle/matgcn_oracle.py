"""CPU oracle for the Multi-ATGCN recurrent graph-convolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``multistgraph_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker /
reported baseline, never as the thing shipped.

It is a from-scratch restatement (plain torch-on-CPU tensor algebra, no
``nn.Module``; float32 by default, float64 on request) of the algorithm in the
reference file ``libcity/model/traffic_flow_prediction/MultiATGCN.py`` ("MA.py"
below).  Every function cites the MA.py lines it follows.  It deliberately keeps
the reference's *per-call* structure (adjacency, Chebyshev stack and per-node
weights are re-derived on every graph-conv call, 96 times per forward) so that
timing it is representative of the reference's CPU path.

Parity pin: ``tests/test_oracle_pin.py`` checks this file against the real
MA.py imported from ``/root/reference`` (when that mount exists) and against the
frozen vectors in ``tests/golden/`` that ``tests/golden/make_golden.py``
generated from the real MA.py.  The reference itself ships no tests or golden
vectors, so those two are the pin.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# a1. scaled Laplacian  (MA.py:15-38)
# --------------------------------------------------------------------------
def scaled_laplacian(adj: np.ndarray, lambda_max: float = 2.0) -> np.ndarray:
    """L~ = (2/lambda) * (I - D^-1/2 A^T D^-1/2) - I with D = row sums of A.

    MA.py:15-23 builds I - (A D^-1/2)^T D^-1/2 (note the transpose, which
    matters for the asymmetric OD view); MA.py:26-38 rescales with the fixed
    ``lambda_max=2`` and ``undirected=False`` the model always passes, and casts
    to float32.  The degree vector is raised to -1/2 in the input dtype and
    infinities (isolated rows) become 0 (MA.py:19-20).
    """
    a = np.asarray(adj)
    deg = a.sum(axis=1)
    with np.errstate(divide="ignore"):
        dis = np.power(deg, -0.5).flatten()
    dis[np.isinf(dis)] = 0.0
    # (A * dis[col]).T * dis[col]  ->  out[j, i] = A[i, j] * dis[j] * dis[i]
    x = (a * dis[None, :]).T * dis[None, :]
    eye = np.eye(a.shape[0], dtype=np.float64)
    lap = eye - x.astype(np.float64)
    out = (2.0 / lambda_max) * lap - eye
    return np.asarray(out.astype(np.float32))


# --------------------------------------------------------------------------
# a2. haversine distances + Gaussian kernel  (MA.py:41-56, 253-261)
# --------------------------------------------------------------------------
def haversine_km(lat1, lng1, lat2, lng2):
    """Great-circle distance in km on a 6371 km sphere (MA.py:41-48)."""
    lat1, lng1, lat2, lng2 = (np.radians(v) for v in (lat1, lng1, lat2, lng2))
    dlat = lat2 - lat1
    dlng = lng2 - lng1
    d = np.sin(dlat * 0.5) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin(dlng * 0.5) ** 2
    return 2 * 6371 * np.arcsin(np.sqrt(d))


def gaussian_kernel_adjacency(dist: np.ndarray, eps: float) -> np.ndarray:
    """exp(-(d/std)^2) with std over all finite entries, small entries zeroed (MA.py:51-56)."""
    finite = dist[~np.isinf(dist)].flatten()
    std = finite.std()
    out = np.exp(-np.square(dist / std))
    out[out < eps] = 0
    return out


_NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?")


def parse_coordinates(coordinate) -> tuple:
    """Return (geo_id array, lon array, lat array) from the ``coordinate`` table.

    MA.py:253-255 splits the "[lon, lat]" strings of the ``coordinates`` column;
    x = first number (longitude), y = second (latitude).  Accepts a pandas
    DataFrame or a dict of columns.
    """
    geo_id = np.asarray(coordinate["geo_id"])
    lon, lat = [], []
    for item in list(coordinate["coordinates"]):
        if isinstance(item, str):
            nums = _NUM.findall(item)
            lon.append(float(nums[0]))
            lat.append(float(nums[1]))
        else:
            lon.append(float(item[0]))
            lat.append(float(item[1]))
    return geo_id, np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)


def distance_view(coordinate) -> np.ndarray:
    """N x N Gaussian-kernel distance adjacency (MA.py:253-261).

    The reference expands all N^2 pairs into a DataFrame and pivots on
    (geo_id, geo_id_1); a pivot sorts both axes, so row/column order is the
    *sorted* geo_id order.  Entry [i, j] is the distance between sorted node i
    and sorted node j (symmetric).
    """
    geo_id, lon, lat = parse_coordinates(coordinate)
    order = np.argsort(geo_id, kind="stable")
    lon, lat = lon[order], lat[order]
    d = haversine_km(lat[:, None], lon[:, None], lat[None, :], lon[None, :])
    return gaussian_kernel_adjacency(d, 0.1)


# --------------------------------------------------------------------------
# a3. static adjacency views and their supports  (MA.py:238-283)
# --------------------------------------------------------------------------
def build_static_supports(config: dict, data_feature: dict) -> Dict[str, object]:
    """Returns {'adj_mx': Tensor used for the SVD init, 'supports': [Tensor N x N, ...]}.

    OD view: adj / diag broadcast over columns, clipped to 1 (MA.py:238-241).
    "cos" view: 1/euclidean distance of static rows, zero distance -> 1, or I
    without static features (MA.py:244-250).  Distance view: MA.py:253-261.
    View selection by ``adjtype`` (MA.py:266-283); each returned support is the
    T_1 of its set (the identity T_0 is implicit).
    """
    n = int(data_feature.get("num_nodes", 1))
    od = torch.tensor(np.asarray(data_feature["adj_mx"]), dtype=torch.float32)
    od = od / torch.diag(od, 0)
    od[od > 1] = 1
    static = data_feature.get("static", None)
    if static is not None:
        s = np.asarray(static, dtype=np.float64)
        diff = s[:, None, :] - s[None, :, :]
        euc = np.sqrt((diff * diff).sum(-1))
        euc[euc == 0] = 1
        cos = torch.tensor(1.0 / euc, dtype=torch.float32)
    else:
        cos = torch.eye(n)
    dis = torch.tensor(distance_view(data_feature["coordinate"]), dtype=torch.float32)

    adjtype = config.get("adjtype", "od")

    def lap(m: Tensor) -> Tensor:
        return torch.tensor(scaled_laplacian(m.numpy()), dtype=torch.float32)

    if adjtype == "multi":
        adj_mx, sup = od, [lap(od), lap(dis), lap(cos)]
    elif adjtype == "od":
        adj_mx, sup = od, [lap(od)]
    elif adjtype == "dist":
        adj_mx, sup = dis, [lap(dis)]
    elif adjtype == "cosine":
        adj_mx, sup = cos, [lap(cos)]
    elif adjtype == "identity":
        adj_mx, sup = torch.eye(n), [torch.eye(n)]
    else:
        raise ValueError("unknown adjtype %r" % (adjtype,))
    return {"adj_mx": adj_mx, "supports": sup, "od": od, "dist": dis, "cos": cos}


# --------------------------------------------------------------------------
# a5-a9. one node-adaptive graph convolution call  (MA.py:76-109)
# --------------------------------------------------------------------------
def agcn_call(x: Tensor, emb: Tensor, vec1: Tensor, vec2: Tensor, static_sup: Sequence[Tensor],
              weights_g: Tensor, weights_pool: Tensor, bias_pool: Tensor,
              cheb_k: int, adjtype: str, adpadj: str) -> Tensor:
    """x [B,N,I] -> [B,N,O], re-deriving adjacency, supports and per-node weights.

    Adaptive view softmax(relu(E E^T)) or softmax(relu(V1 V2)) over dim 1
    (MA.py:80-85); set order [adaptive, static...] and the Chebyshev recurrence
    T_k = 2 T_1 T_{k-1} - T_{k-2} (MA.py:87-101); view weights softmax(weights_g)
    only for ``multi`` (MA.py:102-103); per-node weights and bias from the pools
    (MA.py:104-105); propagation and node-wise contraction (MA.py:106-108).
    """
    n = emb.shape[0]
    eye = torch.eye(n, dtype=x.dtype)
    sets: List[Tensor] = []
    if adpadj == "unidirection":
        sets.append(F.softmax(F.relu(vec1 @ vec2), dim=1))
    elif adpadj == "bidirection":
        sets.append(F.softmax(F.relu(emb @ emb.T), dim=1))
    elif adpadj != "none":
        raise ValueError(adpadj)
    if adpadj == "none":
        sets = [s.to(x.dtype) for s in static_sup]
    elif adjtype == "multi":
        sets = sets + [s.to(x.dtype) for s in static_sup]
    stack = [eye]
    for t1 in sets:
        prev2, prev1 = eye, t1
        stack.append(t1)
        for _ in range(2, cheb_k):
            nxt = (2 * t1) @ prev1 - prev2
            stack.append(nxt)
            prev2, prev1 = prev1, nxt
    sup = torch.stack(stack, dim=0)
    if adjtype == "multi":
        sup = F.softmax(weights_g, dim=0) * sup
    w = torch.einsum("nd,dkio->nkio", emb, weights_pool)
    b = emb @ bias_pool
    xg = torch.einsum("knm,bmc->bknc", sup, x).permute(0, 2, 1, 3)
    return torch.einsum("bnki,nkio->bno", xg, w) + b


def _gru_algebra(x: Tensor, h: Tensor, gate_fn, update_fn, hid: int) -> Tensor:
    """Shared GRU algebra of MA.py:120-128 and MA.py:142-150: z first, r second;
    z gates the candidate input, r is the carry gate."""
    zr = torch.sigmoid(gate_fn(torch.cat((x, h), dim=-1)))
    z, r = torch.split(zr, hid, dim=-1)
    hc = torch.tanh(update_fn(torch.cat((x, z * h), dim=-1)))
    return r * h + (1 - r) * hc


class OracleModel:
    """Functional restatement of ``MultiATGCN`` (MA.py:221-430) over a plain
    name -> tensor dict that uses the reference's state_dict names."""

    def __init__(self, config: dict, data_feature: dict, params: Dict[str, Tensor],
                 dtype: torch.dtype = torch.float32):
        self.cfg = config
        self.df = data_feature
        self.dtype = dtype
        self.p = {k: v.detach().to("cpu", dtype).clone().requires_grad_(v.dtype.is_floating_point)
                  for k, v in params.items()}
        if config.get("node_specific_off", False):  # frozen all-ones embedding, MA.py:350-354
            self.p["node_emb"].requires_grad_(False)
        g = config.get
        self.n = int(data_feature.get("num_nodes", 1))
        self.t_in = g("input_window", 1)
        self.t_out = g("output_window", 1)
        self.hid = g("rnn_units", 64)
        self.layers = g("num_layers", 2)
        self.cheb_k = g("cheb_order", 2)
        self.adjtype = g("adjtype", "od")
        self.adpadj = g("adpadj", "bidirection")
        self.gcn_off = g("gcn_off", False)
        self.fnn_off = g("fnn_off", False)
        self.start_dim = g("start_dim", 0)
        self.end_dim = g("end_dim", 1)
        self.out_dim = self.end_dim - self.start_dim
        self.load_dynamic = g("load_dynamic", False)
        tid, diw = g("add_time_in_day", False), g("add_day_in_week", False)
        # MA.py:313-318
        self.time_dim = 8 if (tid and diw) else (1 if tid else 0)
        self.add_time_in_day = tid
        self.len_c = data_feature.get("len_closeness", 0)
        self.len_p = data_feature.get("len_period", 0)
        self.len_t = data_feature.get("len_trend", 0)
        self.scaler = data_feature.get("scaler")
        # add_static=true (MA.py:286-294, 335-338, 405-409): static node features -> initial hidden state
        st = data_feature.get("static", None)
        self.static = None if st is None else torch.as_tensor(np.asarray(st), dtype=torch.float32)
        self.static_sup = [s.to(dtype) for s in build_static_supports(config, data_feature)["supports"]]
        self.training = False

    # -- parameter helpers ------------------------------------------------
    def grads(self) -> Dict[str, Optional[Tensor]]:
        return {k: v.grad for k, v in self.p.items()}

    def zero_grad(self):
        for v in self.p.values():
            v.grad = None

    def _agcn(self, prefix: str, x: Tensor) -> Tensor:
        p = self.p
        return agcn_call(x, p["node_emb"], p["node_vec1"], p["node_vec2"], self.static_sup,
                         p[prefix + ".weights_g"], p[prefix + ".weights_pool"], p[prefix + ".bias_pool"],
                         self.cheb_k, self.adjtype, self.adpadj)

    def _linear(self, prefix: str, x: Tensor) -> Tensor:
        return x @ self.p[prefix + ".weight"].T + self.p[prefix + ".bias"]

    # -- a10-a12: encoder (MA.py:194-212) -----------------------------------
    def encoder(self, x: Tensor, h0: Tensor) -> Tensor:
        assert x.shape[2] == self.n
        mix = torch.sigmoid(self.p["encoder.weights_gru"])
        cur = x
        for l in range(self.layers):
            h = h0[l]
            outs = []
            cell = "encoder.agru_cells.%d" % l
            for t in range(cur.shape[1]):
                xt = cur[:, t]
                if self.gcn_off:
                    h = _gru_algebra(xt, h, lambda v: self._linear(cell + ".gate", v),
                                     lambda v: self._linear(cell + ".update", v), self.hid)
                else:
                    h = _gru_algebra(xt, h, lambda v: self._agcn(cell + ".gate", v),
                                     lambda v: self._agcn(cell + ".update", v), self.hid)
                    res = "encoder.res_cells.%d" % l
                    rc = _gru_algebra(xt, h, lambda v: self._linear(res + ".gate", v),
                                      lambda v: self._linear(res + ".update", v), self.hid)
                    h = mix[l][t] * h + (1 - mix[l][t]) * rc
                outs.append(h)
            cur = torch.stack(outs, dim=1)
        return cur

    # -- a14: multi-head temporal fusion (MA.py:363-402) --------------------
    def fuse(self, xb: Tensor) -> Tensor:
        src = xb[:, :, :, self.start_dim:self.end_dim]
        wg = F.softmax(self.p["weight_tsg"], dim=0)
        out = 0.0
        c = 0
        if self.len_c > 0:
            b0 = 0
            for _ in range(int(self.len_c / 24)):
                out = out + wg[c] * src[:, b0:b0 + 24] * self.p["weight_ts.%d" % c]
                b0 += 24
                c += 1
        if self.len_p > 0 and self.t_out >= 6:
            b0 = self.len_c
            for _ in range(int(self.len_p / 24)):
                out = out + wg[c] * src[:, b0:b0 + 24] * self.p["weight_ts.%d" % c]
                b0 += 24
                c += 1
        if self.len_t > 0 and self.t_out >= 6:
            b0 = self.len_c + self.len_p  # never advanced: MA.py:388-393
            for _ in range(int(self.len_t / 24)):
                out = out + wg[c] * src[:, b0:b0 + 24] * self.p["weight_ts.%d" % c]
                c += 1
        if self.add_time_in_day:
            out = torch.cat((out, xb[:, 0:self.t_in, :, self.end_dim:self.end_dim + self.time_dim]), dim=-1)
        if self.load_dynamic:
            out = torch.cat((out, xb[:, 0:self.t_in, :, self.end_dim + self.time_dim:]), dim=-1)
        return out

    # -- forward / predict / loss (MA.py:363-430) ---------------------------
    def forward(self, batch) -> Tensor:
        xb = batch["X"].to("cpu", self.dtype)
        fused = self.fuse(xb)
        h0 = torch.zeros(self.layers, xb.shape[0], self.n, self.hid, dtype=self.dtype)
        if self.static is not None:
            # MA.py:405-409: pca_lowrank is called on the float32 features exactly like the reference (it is randomised:
            # parity needs the same RNG state or a patched torch.pca_lowrank), the rest runs in the oracle's dtype
            q = min(self.n, self.p["node_emb"].shape[1]) if "static_initial_gru.embd.weight" not in self.p else \
                self.p["static_initial_gru.embd.weight"].shape[1]
            _, _, v = torch.pca_lowrank(self.static, q=q)
            emb = torch.relu(self._linear("static_initial_gru.embd", (self.static @ v).to(self.dtype)))
            h0 = emb.expand(self.layers, xb.shape[0], -1, -1)
        enc = self.encoder(fused, h0)
        if self.fnn_off:
            enc = enc[:, -1:, :, :]
        enc = F.dropout(enc, p=0.1, training=self.training)
        y = F.conv2d(enc, self.p["end_conv.weight"], self.p["end_conv.bias"])
        return y.squeeze(-1).reshape(-1, self.t_out, self.out_dim, self.n).permute(0, 1, 3, 2)

    predict = forward

    def calculate_loss(self, batch) -> Tensor:
        y_true = batch["y"].to("cpu", self.dtype)
        y_pred = self.forward(batch)
        y_true = self.scaler.inverse_transform(y_true[..., self.start_dim:self.end_dim])
        y_pred = self.scaler.inverse_transform(y_pred)
        return masked_mae(y_pred, y_true, 0.0)


def masked_mae(pred: Tensor, true: Tensor, null_val: float = float("nan"), min_s: float = 1e-4) -> Tensor:
    """libcity/model/loss.py:17-29: |true|<min_s is zeroed (on a copy here; the
    reference does it in place on a temporary), the mask is normalised by its
    mean and NaNs are replaced by 0."""
    true = torch.where(true.abs() < min_s, torch.zeros_like(true), true)
    mask = ~torch.isnan(true) if math.isnan(null_val) else true.ne(null_val)
    mask = mask.to(pred.dtype)
    mask = mask / mask.mean()
    mask = torch.where(torch.isnan(mask), torch.zeros_like(mask), mask)
    loss = (pred - true).abs() * mask
    loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return loss.mean()


class StandardScaler:
    """libcity/utils/normalization.py:62-76."""

    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def transform(self, data):
        return (data - self.mean) / self.std

    def inverse_transform(self, data):
        return data * self.std + self.mean
