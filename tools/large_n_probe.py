"""BASELINE config 5 shape probe (N=8192 census-tract graph): one forward+backward of the drop-in model at a small
per-GPU batch, to show the path runs at the largest graph size and what it costs in time and memory.
usage: python tools/large_n_probe.py [N] [batch] [mode] [micro_batches]   (micro_batches > 0: full train steps through
fused_train_step with gradient accumulation over that many batch slices)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
b = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
micro = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
t0 = time.time()
cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=20, output_window=24, batch_size=b, device=dev)
cfg["matgcn_mode"] = mode
df = make_data_feature(n, seed=0)
torch.manual_seed(0)
model = MultiATGCN(dict(cfg), df).to(dev).train()
print("model built in %.1f s (N=%d)" % (time.time() - t0, n), flush=True)
batch = {k: v.to(dev) for k, v in make_batch(n, b, 24, seed=1).items()}
if micro > 0:
    from multistgraph_b200.train import FusedClipAdam, fused_train_step

    opt = FusedClipAdam(model.parameters(), lr=0.003, max_grad_norm=5.0)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = fused_train_step(model, batch, opt, micro_batches=micro)
        e1.record()
        torch.cuda.synchronize()
        print("train step %d (%d micro-batches of %d): loss %.5f grad-norm %.4e  %.1f ms  (%.2f samples/s)  peak mem %.1f GB"
              % (it, micro, b // micro, float(loss), float(opt.grad_norm), e0.elapsed_time(e1), b / (e0.elapsed_time(e1) * 1e-3),
                 torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
    assert float(loss) == float(loss)
    sys.exit(0)
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.zero_grad()
    loss = model.calculate_loss(batch)
    loss.backward()
    e1.record()
    torch.cuda.synchronize()
    gn = sum(float(p.grad.double().pow(2).sum()) for p in model.parameters() if p.grad is not None) ** 0.5
    print("iter %d: loss %.5f grad-norm %.4e  %.1f ms  (%.2f samples/s)  peak mem %.1f GB"
          % (it, float(loss), gn, e0.elapsed_time(e1), b / (e0.elapsed_time(e1) * 1e-3), torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
assert loss == loss and gn == gn
