"""Kernel-level summary of one warm train step with torch.profiler (CUPTI): where the non-library time goes.
usage: python tools/torch_profile_step.py [mode] [workload] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import workload
from multistgraph_b200.train import FusedClipAdam, fused_train_step

dev = torch.device("cuda:0")
wl = sys.argv[2] if len(sys.argv) > 2 else "baltimore_multi"
cfg, df, batch = workload(wl, seed=0, device=dev, batch=int(sys.argv[3]) if len(sys.argv) > 3 else None)
cfg["matgcn_mode"] = sys.argv[1] if len(sys.argv) > 1 else "bf16"
torch.manual_seed(0)
model = MultiATGCN(dict(cfg), df).to(dev).train()
opt = FusedClipAdam(model.parameters(), lr=0.003, max_grad_norm=5.0)
for _ in range(4):
    fused_train_step(model, batch, opt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        fused_train_step(model, batch, opt)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if t > 0 and e.device_type.name == "CUDA":
        rows.append((t / 3.0, e.count / 3.0, e.key[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("device time per step: %.1f us over %d kernel kinds" % (tot, len(rows)))
for t, c, k in rows[:45]:
    print("%9.1f us  %6.1f x %8.1f us  %5.1f%%  %s" % (t, c, t / max(c, 1e-9), 100 * t / tot, k))
if os.environ.get("DETAIL"):
    # individual copies / stock elementwise kernels of the profiled steps, largest first (what the Python side still launches)
    ev = [e for e in prof.events() if e.device_type.name == "CUDA" and ("Memcpy" in e.name or "at::native" in e.name or "Memset" in e.name)]
    ev.sort(key=lambda e: -(getattr(e, "device_time", 0) or 0))
    print("largest stock copies / elementwise kernels (3 steps):")
    for e in ev[:40]:
        print("%9.1f us  %s" % (getattr(e, "device_time", 0) or 0, e.name[:120]))
