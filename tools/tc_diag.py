"""Layout probe for the tcgen05 GEMM kernel: feeds one-hot operands with exactly representable
values and prints which operand element lands at which output coordinate."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi

DEV = "cuda:0"
lib = _cabi.lib()
st = torch.cuda.current_stream().cuda_stream


def run(a_kc, b_kc, A, B, flags=1):
    M, K = A.shape
    N = B.shape[1]
    pad = lambda v: (v + 3) // 4 * 4
    if a_kc:
        Ad = torch.zeros(M, pad(K)); Ad[:, :K] = A; lda = pad(K)
    else:
        Ad = torch.zeros(K, pad(M)); Ad[:, :M] = A.t(); lda = pad(M)
    if b_kc:
        Bd = torch.zeros(N, pad(K)); Bd[:, :K] = B.t(); ldb = pad(K)
    else:
        Bd = torch.zeros(K, pad(N)); Bd[:, :N] = B; ldb = pad(N)
    Ad, Bd = Ad.to(DEV), Bd.to(DEV)
    C = torch.full((M, N), float("nan"), device=DEV)
    _cabi.check(lib.matgcn_gemm_debug(a_kc, b_kc, M, N, K, Ad.data_ptr(), lda, Bd.data_ptr(), ldb, C.data_ptr(), N, 1, flags, st), "dbg")
    torch.cuda.synchronize()
    return C.cpu()


def summarize(tag, C, ref):
    err = (C - ref).abs().max().item()
    print("%s: max abs err %.3g  nan=%d zeros=%d/%d" % (tag, err, torch.isnan(C).sum().item(), (C == 0).sum().item(), C.numel()))


for (a_kc, b_kc) in [(1, 0), (0, 0), (1, 1)]:
    for K in (8, 32, 64):
        M = N = 128
        # probe B: A one-hot
        A = torch.zeros(M, K)
        for m in range(M):
            A[m, m % K] = 1.0
        B = torch.zeros(K, N)
        for k in range(K):
            for n in range(N):
                B[k, n] = float((k % 8) * 128 + n)
        C = run(a_kc, b_kc, A, B)
        ref = A @ B
        summarize("a_kc=%d b_kc=%d K=%d probeB" % (a_kc, b_kc, K), C, ref)
        if (C - ref).abs().max() > 0 or torch.isnan(C).any():
            print(" row0 C  :", C[0, :16].tolist())
            print(" row0 ref:", ref[0, :16].tolist())
            print(" row1 C  :", C[1, :16].tolist())
            print(" row9 C  :", C[9, :16].tolist(), " ref:", ref[9, :4].tolist())
            print(" row0 C[32:48]:", C[0, 32:48].tolist())
        # probe A: B one-hot
        Bm = torch.zeros(K, N)
        for n in range(N):
            Bm[n % K, n] = 1.0
        Am = torch.zeros(M, K)
        for m in range(M):
            for k in range(K):
                Am[m, k] = float(m * 8 + (k % 8))
        C = run(a_kc, b_kc, Am, Bm)
        ref = Am @ Bm
        summarize("a_kc=%d b_kc=%d K=%d probeA" % (a_kc, b_kc, K), C, ref)
        if (C - ref).abs().max() > 0 or torch.isnan(C).any():
            print(" row0 C  :", C[0, :16].tolist())
            print(" row0 ref:", ref[0, :16].tolist())
            print(" row1 C  :", C[1, :16].tolist(), " ref:", ref[1, :4].tolist())
            print(" col0 C  :", C[:16, 0].tolist())
            print(" col0 ref:", ref[:16, 0].tolist())
            print(" row33 C :", C[33, :8].tolist(), " ref:", ref[33, :8].tolist())
