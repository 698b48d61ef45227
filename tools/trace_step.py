import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import workload
dev = torch.device("cuda:0")
cfg, df, batch = workload("baltimore_multi", seed=0, device=dev)
cfg["matgcn_mode"] = sys.argv[1] if len(sys.argv) > 1 else "tf32"
torch.manual_seed(0)
model = MultiATGCN(dict(cfg), df).to(dev).train()
for i in range(3):
    if i == 2:
        print("==== traced step ====", file=sys.stderr, flush=True)
    model.zero_grad()
    model.calculate_loss(batch).backward()
    torch.cuda.synchronize()
