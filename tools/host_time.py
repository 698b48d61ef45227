"""Host time to enqueue one train step (no synchronisation inside the loop), persistent kernels on / off, and where it goes."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import workload
from multistgraph_b200.train import FusedClipAdam, fused_train_step

dev = torch.device("cuda:0")
cfg, df, batch = workload("baltimore_multi", seed=0, device=dev)
cfg["matgcn_mode"] = "bf16"
torch.manual_seed(0)
model = MultiATGCN(dict(cfg), df).to(dev).train()
opt = FusedClipAdam(model.parameters(), lr=0.003, max_grad_norm=5.0)
lib = _cabi.lib()
for rec in (1, 0):
    lib.matgcn_set_recurrent_kernel(rec)
    for _ in range(3):
        fused_train_step(model, batch, opt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fused_train_step(model, batch, opt)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("persistent=%d: host enqueue %.2f ms per step, wall %.2f ms per step" % (rec, (t1 - t0) * 200, (t2 - t0) * 200))
lib.matgcn_set_recurrent_kernel(1)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    fused_train_step(model, batch, opt)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
