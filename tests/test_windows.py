"""CPU tests of the oracle for SURVEY.md 8f rows f1 (clip + Adam) and f2 (batch assembly):
the numpy restatements are pinned against golden vectors generated from the real reference class, against the
reference itself when /root/reference is mounted, and against torch's own clip_grad_norm_ + Adam."""
import logging
import os
import sys

import numpy as np
import pytest
import torch

from oracle import window_oracle as wo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "windows", "windows_small.npz")
sys.path.insert(0, os.path.join(HERE, "golden", "windows"))

CASES = {
    "shipped_like": (24 * 30, 5, 2, 24, 24, 2, 1, 1, 7, 28, 1, 24),
    "short_out": (24 * 12, 3, 3, 24, 3, 2, 1, 0, 7, 7, 1, 24),
    "closeness_only": (200, 7, 1, 12, 6, 3, 0, 0, 1, 7, 1, 24),
    "half_hour": (48 * 10, 2, 2, 24, 12, 1, 2, 1, 1, 3, 2, 24),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_window_oracle_matches_golden_from_reference(name):
    g = np.load(GOLD)
    T, N, F, iw, ow, lc, lp, lt, ip, it, pph, hed = CASES[name]
    df = g[name + "/series"]
    assert df.shape == (T, N, F)
    x, y, starts = wo.generate_input_data(df, iw, ow, lc, lp, lt, ip, it, pph, hed)
    assert x.shape[0] == int(g[name + "/n_samples"]) == len(starts)
    pick = g[name + "/pick"]
    assert np.array_equal(x[pick], g[name + "/x_pick"])          # copies: bit-exact
    assert np.array_equal(y[pick], g[name + "/y_pick"])
    assert np.array_equal(x.astype(np.float64).sum(axis=(1, 2, 3)), g[name + "/x_sum"])
    assert np.array_equal(y.astype(np.float64).sum(axis=(1, 2, 3)), g[name + "/y_sum"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/libcity"), reason="reference mount not present")
def test_window_oracle_matches_live_reference():
    from make_windows_golden import reference_windows

    rng = np.random.default_rng(7)
    df = rng.standard_normal((24 * 9, 4, 2)).astype(np.float32)
    args = (24, 6, 3, 1, 1, 1, 7, 1, 24)
    xr, yr = reference_windows(df, *args)
    x, y, _ = wo.generate_input_data(df, *args)
    assert np.array_equal(x, xr) and np.array_equal(y, yr)


def test_window_oracle_rejects_out_of_range_samples():
    # no sample may reach before the series start or past its end (mth_dataset.py:45-46, 52-58)
    assert wo.sample_indices(100, 10, 24, 1, 0, 0, 1, 7, 1, 24) is None      # closeness would start at -14
    assert wo.sample_indices(100, 77, 24, 1, 0, 0, 1, 7, 1, 24) is None      # label window ends at 101
    assert wo.sample_indices(100, 76, 24, 1, 0, 0, 1, 7, 1, 24) == ([(52, 76)], [], [])


@pytest.mark.parametrize("wd,max_norm", [(0.0, 5.0), (0.0, None), (1e-3, 0.5)])
def test_adam_oracle_matches_torch_clip_and_adam(wd, max_norm):
    """Pins oracle.adam_clip_reference on torch itself: clip_grad_norm_ (executor:420-421) + Adam (executor:146-147)."""
    torch.manual_seed(0)
    shapes = [(7, 5), (13,), (3, 4, 2)]
    params = [torch.nn.Parameter(torch.randn(s, dtype=torch.float64)) for s in shapes]
    opt = torch.optim.Adam(params, lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    p = np.concatenate([q.detach().numpy().ravel() for q in params])
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for step in range(1, 5):
        grads = [torch.randn(s, dtype=torch.float64) * (3.0 if step % 2 else 0.1) for s in shapes]
        for q, g in zip(params, grads):
            q.grad = g.clone()
        if max_norm is not None:
            total_t = torch.nn.utils.clip_grad_norm_(params, max_norm)
        opt.step()
        g_flat = np.concatenate([g.numpy().ravel() for g in grads])
        p, g_after, m, v, total = wo.adam_clip_reference(p, g_flat, m, v, step, 0.01, 0.9, 0.999, 1e-8, wd, max_norm)
        want = np.concatenate([q.detach().numpy().ravel() for q in params])
        assert np.allclose(p, want, rtol=1e-12, atol=1e-14)
        if max_norm is not None:
            assert abs(total - float(total_t)) < 1e-12 * max(1.0, total)
            assert np.allclose(g_after, np.concatenate([q.grad.numpy().ravel() for q in params]), rtol=1e-12)


def test_train_side_helpers_have_no_cpu_fallback():
    """The product path of the 8f rows fails loudly without CUDA tensors (no silent PyTorch / oracle route)."""
    from multistgraph_b200 import ops
    from multistgraph_b200._cabi import MatgcnError
    from multistgraph_b200.train import DeviceWindowBank, FusedClipAdam

    with pytest.raises(MatgcnError):
        FusedClipAdam([torch.nn.Parameter(torch.zeros(3))])
    with pytest.raises(MatgcnError):
        DeviceWindowBank(torch.zeros(100, 3, 2), 24, 24, 1, 0, 0)
    with pytest.raises(MatgcnError):
        ops.output_head(torch.zeros(2, 3, 4, 64), torch.zeros(5, 2, 64), torch.zeros(5))
