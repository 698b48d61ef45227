"""Prints the downstream numbers of tests/test_gpu_acceptance.py for all three engines and a longer run.
usage: python tools/train_acceptance.py [steps] [N]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.acceptance import train_and_evaluate

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n = int(sys.argv[2]) if len(sys.argv) > 2 else 237
out = {}
for mode in ("exact", "tf32", "bf16"):
    before, after, losses = train_and_evaluate(mode, n_nodes=n, steps=steps)
    out[mode] = {"untrained_MAE": before["MAE"], "MAE": after["MAE"], "RMSE": after["RMSE"], "first_loss": losses[0], "last_loss": losses[-1]}
    print(mode, json.dumps(out[mode]), flush=True)
before, after, losses = train_and_evaluate("exact", n_nodes=n, steps=steps, order_seed=77)
out["exact, other batch order / dropout seed"] = {"MAE": after["MAE"], "RMSE": after["RMSE"], "last_loss": losses[-1]}
print("exact, other batch order / dropout seed", json.dumps(out["exact, other batch order / dropout seed"]), flush=True)
ref = out["exact"]
for mode in ("tf32", "bf16", "exact, other batch order / dropout seed"):
    print(mode, "relative to exact:", {k: ["%.2f%%" % (100 * (a - b) / b) for a, b in zip(out[mode][k], ref[k])] for k in ("MAE", "RMSE")})
