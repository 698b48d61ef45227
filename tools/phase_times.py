"""CUDA-event breakdown of one train step (warm): fusion / encoder layers fwd / head+loss / backward pieces / optimizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import ops, _cabi
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.dp import FlatGradBucket
from multistgraph_b200.synthetic import workload

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "baltimore_multi"
mode = sys.argv[2] if len(sys.argv) > 2 else "tf32"
bsz = int(sys.argv[3]) if len(sys.argv) > 3 else None
cfg, df, batch = workload(name, seed=0, batch=bsz, device=dev)
cfg["matgcn_mode"] = mode
torch.manual_seed(0)
model = MultiATGCN(dict(cfg), df).to(dev).train()
bucket = FlatGradBucket(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=0.003)

marks = []
def mark(tag):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((tag, e))

orig_layer = ops.encoder_layer
class Wrap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tag):
        ctx.tag = tag
        mark("fwd:" + tag)
        return x.view_as(x)
    @staticmethod
    def backward(ctx, g):
        mark("bwd:" + ctx.tag)
        return g, None
layer_idx = [0]
def timed_layer(*a):
    i = layer_idx[0]; layer_idx[0] += 1
    x = Wrap.apply(a[0], "layer%d_in" % i)
    y = orig_layer(x, *a[1:])
    return Wrap.apply(y, "layer%d_out" % i)
ops.encoder_layer = timed_layer

def step():
    marks.clear(); layer_idx[0] = 0
    mark("start")
    bucket.zero()
    loss = model.calculate_loss(batch)
    mark("fwd:done")
    loss.backward()
    mark("bwd:done")
    bucket.clip_grad_norm_(5.0); opt.step()
    mark("opt:done")
    torch.cuda.synchronize()
    return [(t, marks[0][1].elapsed_time(e)) for t, e in marks]

for _ in range(3):
    step()
acc = {}
R = 5
for _ in range(R):
    for t, ms in step():
        acc[t] = acc.get(t, 0.0) + ms / R
prev = 0.0
for t, ms in sorted(acc.items(), key=lambda kv: kv[1]):
    print("%-18s at %8.3f ms   (+%.3f)" % (t, ms, ms - prev)); prev = ms
print("peak memory GB", torch.cuda.max_memory_allocated() / 2**30)
