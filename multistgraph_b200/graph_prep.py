"""Init-time graph preparation (host side, numpy): the static adjacency views and
their scaled Laplacians.

Replaces, for the drop-in model, the scipy-sparse / pandas code of the reference:

* ``calculate_normalized_laplacian`` / ``calculate_scaled_laplacian``
  (MultiATGCN.py:15-38) -> :func:`scaled_laplacian`
* ``haversine_array`` / ``calculate_adjacency_matrix_dist`` and the N^2-row pandas
  concat + pivot (MultiATGCN.py:41-56, 253-261) -> :func:`distance_adjacency`
  (broadcast numpy: seconds instead of minutes at N=8192)
* the OD / "cos" / view-selection logic of ``MultiATGCN.__init__``
  (MultiATGCN.py:238-283) -> :func:`static_views`

This runs once per model construction and is not on the per-step path.
"""
from __future__ import annotations

import json
from typing import Dict, List

import numpy as np


def scaled_laplacian(adj: np.ndarray) -> np.ndarray:
    """``-D^-1/2 A^T D^-1/2`` with D the row sums of ``adj``.

    With the fixed ``lambda_max=2`` and ``undirected=False`` the reference passes,
    ``(2/lambda)(I - D^-1/2 A^T D^-1/2) - I`` collapses to this.
    Degree powers are taken in the input dtype like the reference (float32 in ->
    float32 scale factors), rows with zero degree contribute zeros.
    """
    a = np.asarray(adj)
    d = a.sum(axis=1)
    with np.errstate(divide="ignore"):
        s = np.power(d, -0.5)
    s = np.where(np.isinf(s), np.zeros_like(s), s)
    out = ((a * s[np.newaxis, :]).T * s[np.newaxis, :])
    return (-out).astype(np.float32)


def _lon_lat(coordinate):
    ids = np.asarray(coordinate["geo_id"])
    lon = np.empty(len(ids), dtype=np.float64)
    lat = np.empty(len(ids), dtype=np.float64)
    for i, c in enumerate(list(coordinate["coordinates"])):
        pair = json.loads(c) if isinstance(c, str) else c
        lon[i], lat[i] = float(pair[0]), float(pair[1])
    return ids, lon, lat


def distance_adjacency(coordinate, eps: float = 0.1, block: int = 2048) -> np.ndarray:
    """Gaussian-kernel great-circle adjacency, nodes ordered by sorted ``geo_id``
    (the order the reference's pivot produces).  Computed in row blocks so the
    N=8192 case never holds more than ``block`` x N temporaries."""
    ids, lon, lat = _lon_lat(coordinate)
    order = np.argsort(ids, kind="stable")
    lon, lat = np.radians(lon[order]), np.radians(lat[order])
    n = len(ids)
    dist = np.empty((n, n), dtype=np.float64)
    cos_lat = np.cos(lat)
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        dlat = lat[None, :] - lat[r0:r1, None]
        dlon = lon[None, :] - lon[r0:r1, None]
        h = np.sin(dlat * 0.5) ** 2 + cos_lat[r0:r1, None] * cos_lat[None, :] * np.sin(dlon * 0.5) ** 2
        dist[r0:r1] = 2 * 6371 * np.arcsin(np.sqrt(h))
    sigma = dist[np.isfinite(dist)].std()
    out = np.exp(-np.square(dist / sigma))
    out[out < eps] = 0
    return out


def od_view(adj_mx) -> np.ndarray:
    """OD counts divided by the destination's diagonal (broadcast over columns),
    clipped to 1 (MultiATGCN.py:238-241)."""
    a = np.asarray(adj_mx, dtype=np.float32)
    a = a / np.diag(a)[np.newaxis, :]
    a[a > 1] = 1
    return a.astype(np.float32)


def similarity_view(static, n: int) -> np.ndarray:
    """1 / euclidean distance between static feature rows (0 -> 1), or I (MultiATGCN.py:244-250).
    Distances come from explicit differences (scipy ``cdist``, the reference's own call, MA.py:246): a Gram expansion
    ``|a|^2 + |b|^2 - 2ab`` cancels badly for duplicated or nearly equal rows of unnormalised features, and then the
    ``== 0 -> 1`` rule does not fire where the reference's does."""
    if static is None:
        return np.eye(n, dtype=np.float32)
    from scipy.spatial.distance import cdist

    s = np.asarray(static, dtype=np.float64)
    out = np.empty((n, n), dtype=np.float32)
    step = max(1, (1 << 24) // max(1, n))          # row blocks of ~16 M distances
    for r0 in range(0, n, step):
        d = cdist(s[r0:r0 + step], s)
        d[d == 0] = 1
        out[r0:r0 + step] = (1.0 / d).astype(np.float32)
    return out


def static_views(adjtype: str, data_feature: dict) -> Dict[str, object]:
    """-> {'adj_mx': matrix the SVD init of node_vec1/2 uses, 'laplacians': [T_1 per static set]}
    in the reference's set order [od, dist, cos] (MultiATGCN.py:266-283)."""
    n = int(data_feature.get("num_nodes", 1))
    od = od_view(data_feature["adj_mx"])
    cos = similarity_view(data_feature.get("static", None), n)
    dis = distance_adjacency(data_feature["coordinate"]).astype(np.float32)
    eye = np.eye(n, dtype=np.float32)
    if adjtype == "multi":
        adj, laps = od, [scaled_laplacian(od), scaled_laplacian(dis), scaled_laplacian(cos)]
    elif adjtype == "od":
        adj, laps = od, [scaled_laplacian(od)]
    elif adjtype == "dist":
        adj, laps = dis, [scaled_laplacian(dis)]
    elif adjtype == "cosine":
        adj, laps = cos, [scaled_laplacian(cos)]
    elif adjtype == "identity":
        adj, laps = eye, [eye]
    else:
        raise ValueError("adjtype must be one of multi/od/dist/cosine/identity, got %r" % (adjtype,))
    return {"adj_mx": adj, "laplacians": laps}


def chebyshev_terms(t1: np.ndarray, cheb_k: int) -> List[np.ndarray]:
    """[T_1, ..., T_{cheb_k-1}] with T_k = 2 T_1 T_{k-1} - T_{k-2}, T_0 = I
    (MultiATGCN.py:98-100).  For cheb_k <= 2 this is just [T_1] - including the
    cheb_order=1 quirk where T_1 is still appended."""
    n = t1.shape[0]
    terms = [np.eye(n, dtype=np.float32), t1.astype(np.float32)]
    for _ in range(2, cheb_k):
        terms.append((2.0 * t1) @ terms[-1] - terms[-2])
    return [t.astype(np.float32) for t in terms[1:]]
