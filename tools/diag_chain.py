"""Where a tensor-core launch spends its time: per-tile timeline of CTA 0 (tf32 and bf16 propagation) and
back-to-back launch chains (tiny GEMM = launch floor of the kernel structure; propagation warm in L2)."""
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
M16, X16 = M.bfloat16(), X.bfloat16()


def prop32():
    _cabi.check(lib.matgcn_propagate_fwd(M.data_ptr(), kp, n, ldm, X.data_ptr(), cols, P.data_ptr(), 1, st), "p")


def prop16():
    _cabi.check(lib.matgcn_propagate_fwd_bf16(M16.data_ptr(), kp, n, ldm, X16.data_ptr(), cols, P.data_ptr(), st), "p16")


A = torch.randn(128, 32, device=dev)
B = torch.randn(32, 128, device=dev)
C = torch.empty(128, 128, device=dev)


def tiny():
    _cabi.check(lib.matgcn_gemm_debug(1, 0, 128, 128, 32, A.data_ptr(), 32, B.data_ptr(), 128, C.data_ptr(), 128, 1, 1, st), "tiny")


def chain(fn, iters, tag):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    print("%-28s chain of %3d: %.2f us per launch" % (tag, iters, e0.elapsed_time(e1) * 1e3 / iters), flush=True)


names = ["prod_start", "mma_waitbeg", "mma_start", "mma_commit", "epi_waitbeg", "epi_start", "epi_end"]


def timeline(fn, tag, mode=0):
    lib.matgcn_debug_set_mode(mode)
    buf = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.matgcn_debug_set_timeline(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.matgcn_debug_set_timeline(None)
    lib.matgcn_debug_set_mode(0)
    b = buf.cpu().view(-1, 8)
    t0 = b[0, 0].item()
    print("timeline %s (mode %d), cycles since producer start of tile 0:" % (tag, mode))
    for i in range(4):
        if b[i, 0].item() == 0:
            break
        print("  tile %d: " % i + "  ".join("%s=%d" % (nm, b[i, j].item() - t0) for j, nm in enumerate(names)))


print("PDL env:", os.environ.get("MATGCN_PDL"))
chain(tiny, 200, "tiny 128x128x32")
chain(prop32, 50, "prop tf32 403 (warm L2)")
chain(prop16, 50, "prop bf16 403 (warm L2)")
lib.matgcn_debug_set_mode(1)
chain(prop32, 50, "prop tf32 no-store")
chain(prop16, 50, "prop bf16 no-store")
lib.matgcn_debug_set_mode(0)
timeline(prop32, "prop tf32")
timeline(prop16, "prop bf16")
timeline(prop16, "prop bf16 no-store", 1)
