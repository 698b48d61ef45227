"""Generates tests/golden/windows/windows_small.npz from the REAL reference class
(libcity/data/dataset/dataset_subclass/mth_dataset.py, MTHDataset._generate_input_data) on a small seeded series.
Run in the build container only (needs /root/reference):  python tests/golden/windows/make_windows_golden.py"""
import logging
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from libcity.data.dataset.dataset_subclass.mth_dataset import MTHDataset  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_windows(df, input_window, output_window, len_closeness, len_period, len_trend, interval_period,
                      interval_trend, points_per_hour, hour_each_day):
    ds = MTHDataset.__new__(MTHDataset)   # bypass the file-loading constructor; only the window logic is exercised
    ds.points_per_hour = points_per_hour
    ds.len_closeness, ds.len_period, ds.len_trend = len_closeness, len_period, len_trend
    ds.interval_period, ds.interval_trend = interval_period, interval_trend
    ds.hour_each_day = hour_each_day
    ds.input_window, ds.output_window = input_window, output_window
    ds._logger = logging.getLogger("golden")
    return ds._generate_input_data(df)


CASES = {
    # name: (T_total, N, F, input_window, output_window, len_c, len_p, len_t, interval_p, interval_t, pph, hours)
    "shipped_like": (24 * 30, 5, 2, 24, 24, 2, 1, 1, 7, 28, 1, 24),   # config_user.json heads 2/1/1, 7 d / 28 d, shortened series
    "short_out": (24 * 12, 3, 3, 24, 3, 2, 1, 0, 7, 7, 1, 24),
    "closeness_only": (200, 7, 1, 12, 6, 3, 0, 0, 1, 7, 1, 24),
    "half_hour": (48 * 10, 2, 2, 24, 12, 1, 2, 1, 1, 3, 2, 24),
}

if __name__ == "__main__":
    out = {}
    rng = np.random.default_rng(0)
    for name, (T, N, F, iw, ow, lc, lp, lt, ip, it, pph, hed) in CASES.items():
        df = rng.standard_normal((T, N, F)).astype(np.float32)
        x, y = reference_windows(df, iw, ow, lc, lp, lt, ip, it, pph, hed)
        # keep the fixture small: the series, the sample count and a checksum + a few whole samples
        pick = np.unique(np.linspace(0, x.shape[0] - 1, 5).astype(int))
        out[name + "/series"] = df
        out[name + "/n_samples"] = np.int64(x.shape[0])
        out[name + "/pick"] = pick
        out[name + "/x_pick"] = x[pick]
        out[name + "/y_pick"] = y[pick]
        out[name + "/x_sum"] = x.astype(np.float64).sum(axis=(1, 2, 3))
        out[name + "/y_sum"] = y.astype(np.float64).sum(axis=(1, 2, 3))
        print(name, x.shape, y.shape)
    np.savez_compressed(os.path.join(HERE, "windows_small.npz"), **out)
