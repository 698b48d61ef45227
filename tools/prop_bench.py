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
    M16, X16 = M.bfloat16(), X.bfloat16()
    P16 = torch.empty(kp, n, cols, device=dev, dtype=torch.bfloat16)
    modes = (3,) if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else (4,) if (len(sys.argv) > 2 and sys.argv[2] == "twin") else (1, 0)
    for flags in modes:
        def run(flags=flags):
            if flags == 4:   # the launch of a bf16-mode step: bf16 in, bf16 twin out
                _cabi.check(lib.matgcn_propagate_fwd_bf16_twin(M16.data_ptr(), kp, n, ldm, X16.data_ptr(), cols, P16.data_ptr(), st), "twin")
            elif flags == 3:
                _cabi.check(lib.matgcn_propagate_fwd_bf16(M16.data_ptr(), kp, n, ldm, X16.data_ptr(), cols, P.data_ptr(), st), "p16")
            else:
                _cabi.check(lib.matgcn_propagate_fwd(M.data_ptr(), kp, n, ldm, X.data_ptr(), cols, P.data_ptr(), flags, st), "p")
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            run()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * kp * n * n * cols
        print("Kp=%d N=%d cols=%d flags=%d: %.1f us  %.1f TFLOP/s" % (kp, n, cols, flags, ms * 1e3, fl / ms / 1e9), flush=True)
