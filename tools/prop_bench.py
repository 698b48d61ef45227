"""Times the support-propagation kernel (both engines) at several shapes; used under ncu for the per-kernel profile."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi

lib = _cabi.lib()
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
shapes = [(4, 403, 4096), (4, 883, 16384), (4, 2048, 4096)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for kp, n, cols in shapes:
    ldm = (n + 7) // 8 * 8
    M = torch.randn(kp, n, ldm, device=dev) * 0.05
    X = torch.randn(n, cols, device=dev)
    P = torch.empty(kp, n, cols, device=dev)
    for flags in (1, 0):
        for _ in range(2):
            _cabi.check(lib.matgcn_propagate_fwd(M.data_ptr(), kp, n, ldm, X.data_ptr(), cols, P.data_ptr(), flags, st), "p")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            _cabi.check(lib.matgcn_propagate_fwd(M.data_ptr(), kp, n, ldm, X.data_ptr(), cols, P.data_ptr(), flags, st), "p")
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * kp * n * n * cols
        print("Kp=%d N=%d cols=%d flags=%d: %.1f us  %.1f TFLOP/s" % (kp, n, cols, flags, ms * 1e3, fl / ms / 1e9), flush=True)
