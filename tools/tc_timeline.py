"""Per-tile timeline of CTA 0 of the tcgen05 kernel (producer / MMA / epilogue stamps, in SM cycles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi
lib = _cabi.lib()
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
kp, n, cols = 4, 403, 4096
ldm = (n + 7) // 8 * 8
M = torch.randn(kp, n, ldm, device=dev) * 0.05
X = torch.randn(n, cols, device=dev)
P = torch.empty(kp, n, cols, device=dev)
for _ in range(3):
    lib.matgcn_propagate_fwd(M.data_ptr(), kp, n, ldm, X.data_ptr(), cols, P.data_ptr(), 1, st)
torch.cuda.synchronize()
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.matgcn_debug_set_mode(mode)
print("debug mode", mode)
buf = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
lib.matgcn_debug_set_timeline(buf.data_ptr())
lib.matgcn_propagate_fwd(M.data_ptr(), kp, n, ldm, X.data_ptr(), cols, P.data_ptr(), 1, st)
torch.cuda.synchronize()
lib.matgcn_debug_set_timeline(None)
b = buf.cpu().view(-1, 8)
t0 = b[0, 0].item()
names = ["prod_start", "mma_waitbeg", "mma_start", "mma_commit", "epi_waitbeg", "epi_start", "epi_end"]
for i in range(4):
    if b[i, 0].item() == 0:
        break
    print("tile %d: " % i + "  ".join("%s=%d" % (nm, b[i, j].item() - t0) for j, nm in enumerate(names)))
