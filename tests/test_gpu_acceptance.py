"""Downstream acceptance of the fast modes (north star: "bf16/TF32 tensor-core mode: a stated looser bound plus unchanged
MAE/RMSE on the DC test split").  The same seeded training run - DC shape (N = 237, 24 -> 3), executor recipe (Adam 0.003,
clip 5, dropout 0.1) - is made by the exact engine (fp32 FFMA, 1e-4 parity) and by the bf16 engine bench.py quotes, and
both are evaluated on the held-out last fifth of the series with the evaluator's MAE@i / RMSE@i.

Stated bound: every MAE@i / RMSE@i of the bf16 run within 1 % of the exact run's, and training must have learned something
(test MAE at least 5x below the untrained model's).  For scale: two exact-mode runs that differ only in batch order and dropout
seed differ by up to 0.4 % (MAE) / 0.6 % (RMSE) on this split, the bf16 run by 0.1 % / 0.02 % from the exact run that shares its
batches (tools/train_acceptance.py, profiles/r2_train_acceptance.txt) - below what training noise alone produces."""
import pytest
import torch

from tests.acceptance import train_and_evaluate

pytestmark = pytest.mark.gpu


def test_bf16_training_run_matches_exact_mode_on_the_test_split():
    b0, exact, l0 = train_and_evaluate("exact", steps=400)
    b1, fast, l1 = train_and_evaluate("bf16", steps=400)
    print("[acceptance] untrained MAE@3 %.3f | exact: MAE %s RMSE %s | bf16: MAE %s RMSE %s | last train loss %.4f / %.4f"
          % (b0["MAE"][-1], ["%.3f" % v for v in exact["MAE"]], ["%.3f" % v for v in exact["RMSE"]],
             ["%.3f" % v for v in fast["MAE"]], ["%.3f" % v for v in fast["RMSE"]], l0[-1], l1[-1]))
    assert abs(b0["MAE"][-1] - b1["MAE"][-1]) / b0["MAE"][-1] < 2e-3, "the untrained models already disagree"
    assert exact["MAE"][-1] < 0.2 * b0["MAE"][-1], "training did not reduce the test error"
    for name in ("MAE", "RMSE"):
        for i, (a, b) in enumerate(zip(fast[name], exact[name])):
            assert abs(a - b) / b < 0.01, "%s@%d: bf16 %.4f vs exact %.4f" % (name, i + 1, a, b)
    assert all(torch.isfinite(torch.tensor(l1)))
