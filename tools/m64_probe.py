"""Probe of the M=64 tcgen05 accumulator layout: raw dump of all 128 TMEM lanes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi
lib = _cabi.lib()
dev = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
M, N, K = 64, 128, 32
A = torch.zeros(M, K)
for m in range(M):
    A[m, :] = 0.0
    A[m, m % K] = float(m + 1)          # row m -> value m+1 at k = m % 32
B = torch.zeros(K, N)
for n in range(N):
    B[n % K, n] = 1.0                    # C[m, n] = A[m, n % 32] = (m+1) if n%32 == m%32
ref = A @ B
Ad, Bd = A.to(dev), B.to(dev)
for mode, label in [(8, "m64, hypothesised mapping"), (8 | 4, "m64 raw lane dump"), (0, "m128")]:
    lib.matgcn_debug_set_mode(mode)
    rows = 128 if mode & 4 else M
    C = torch.full((128, N), -7.0, device=dev)
    _cabi.check(lib.matgcn_gemm_debug(1, 0, M, N, K, Ad.data_ptr(), K, Bd.data_ptr(), N, C.data_ptr(), N, 1, 1, st), "dbg")
    torch.cuda.synchronize()
    C = C.cpu()
    print("==", label, " max err vs ref (first 64 rows):", (C[:M] - ref).abs().max().item())
    if mode & 4 or (C[:M] - ref).abs().max() > 0:
        # for each output row of the dump, which logical row does it hold?  (value = m+1 at column m%32)
        for r in range(0, 128, 1):
            nz = C[r].nonzero().flatten().tolist()
            vals = sorted(set(C[r][nz].tolist())) if nz else []
            if r < 8 or r % 16 == 0 or r % 16 == 15:
                print("  dump row %3d: nonzero cols %s... values %s" % (r, nz[:4], vals[:4]))
lib.matgcn_debug_set_mode(0)
