"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the batch-assembly row f2 of SURVEY.md section 8.

A plain numpy restatement, sample by sample, of how the reference cuts model inputs out of one
``[T_total, N, F]`` series:

* ``search_data``          follows ``MTHDataset._search_data``          (libcity/data/dataset/dataset_subclass/mth_dataset.py:31-60)
* ``sample_indices``       follows ``MTHDataset._get_sample_indices``   (mth_dataset.py:62-110)
* ``generate_input_data``  follows ``MTHDataset._generate_input_data``  (mth_dataset.py:112-158)

Pinned by ``tests/test_windows.py`` against the real class imported from ``/root/reference`` (when the mount
exists) and against ``tests/golden/windows/windows_small.npz`` generated from it by ``tests/golden/windows/make_windows_golden.py``.
Only tests / smoke / the bench's CPU leg may import this module; the product path is
``multistgraph_b200.train.DeviceWindowBank`` -> ``matgcn_assemble_windows`` (CUDA), which has no CPU fallback.
"""
from __future__ import annotations

import numpy as np


def search_data(sequence_length, label_start_idx, num_for_predict, num_of_depend, units, points_per_hour):
    """mth_dataset.py:31-60: the num_of_depend segments [start, start+num_for_predict) that lie
    int(points_per_hour * units * i) slices before the label start, oldest first; None if any is out of range."""
    if label_start_idx + num_for_predict > sequence_length:      # :45-46
        return None
    x_idx = []
    for i in range(1, num_of_depend + 1):                        # :48-56
        start_idx = label_start_idx - int(points_per_hour * units * i)
        end_idx = start_idx + num_for_predict
        if start_idx >= 0:
            x_idx.append((start_idx, end_idx))
        else:
            return None
    return x_idx[::-1]                                           # :59 oldest -> newest


def sample_indices(seq_len, label_start_idx, input_window, len_closeness, len_period, len_trend,
                   interval_period, interval_trend, points_per_hour, hour_each_day):
    """mth_dataset.py:62-110 reduced to index arithmetic: returns (closeness, period, trend) segment lists or None."""
    if label_start_idx + input_window > seq_len:                 # :78-79
        return None
    trend = period = closeness = []
    if len_trend > 0:                                            # :81-87
        trend = search_data(seq_len, label_start_idx, input_window, len_trend, interval_trend * hour_each_day, points_per_hour)
        if not trend:
            return None
    if len_period > 0:                                           # :89-95
        period = search_data(seq_len, label_start_idx, input_window, len_period, interval_period * hour_each_day, points_per_hour)
        if not period:
            return None
    if len_closeness > 0:                                        # :97-103
        closeness = search_data(seq_len, label_start_idx, input_window, len_closeness, input_window / points_per_hour, points_per_hour)
        if not closeness:
            return None
    return closeness, period, trend


def generate_input_data(df, input_window, output_window, len_closeness, len_period, len_trend,
                        interval_period=1, interval_trend=7, points_per_hour=1, hour_each_day=24):
    """mth_dataset.py:112-158: every valid label start in order; sources = cat(closeness, period, trend) on the
    time axis (:140-153), targets = df[s : s + output_window] (:105).  Also returns the label starts."""
    xs, ys, starts = [], [], []
    for idx in range(df.shape[0]):
        seg = sample_indices(df.shape[0], idx, input_window, len_closeness, len_period, len_trend,
                             interval_period, interval_trend, points_per_hour, hour_each_day)
        if seg is None:
            continue
        closeness, period, trend = seg
        parts = [df[i:j] for i, j in closeness] + [df[i:j] for i, j in period] + [df[i:j] for i, j in trend]
        xs.append(np.concatenate(parts, axis=0))
        ys.append(df[idx: idx + output_window])
        starts.append(idx)
    return np.stack(xs, axis=0), np.stack(ys, axis=0), np.asarray(starts, dtype=np.int64)


def adam_clip_reference(param, grad, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale=1.0):
    """Row f1: clip_grad_norm_ (traffic_state_executor.py:420-421) then torch.optim.Adam.step (executor:146-147), restated
    in float64 numpy from torch's documented single-tensor update.  Returns (param, grad, exp_avg, exp_avg_sq, total_norm)."""
    g = grad.astype(np.float64) * grad_scale
    total = float(np.sqrt(np.sum(g * g)))
    if max_norm is not None and max_norm > 0:
        g = g * min(max_norm / (total + 1e-6), 1.0)
    gd = g + weight_decay * param if weight_decay else g
    m = exp_avg + (gd - exp_avg) * (1.0 - beta1)
    v = beta2 * exp_avg_sq + (1.0 - beta2) * gd * gd
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = param - (lr / bc1) * (m / denom)
    return p, g, m, v, total
