"""One encoder-layer fwd+bwd at the Baltimore shape (T=2) - a short target for ncu --set full captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi
lib = _cabi.lib()
dev = "cuda:0"
T, N, B, Cin, H, Kp, n_adp = 24, 403, 64, 2, 64, 4, 1
K, I = Kp + 1, Cin + H
ldm = (N + 7) // 8 * 8
g = torch.Generator().manual_seed(0)
R = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
x, M = R(T, N, B, Cin), R(Kp, N, ldm, sc=0.05)
Wg, Wu = R(N, K, I, 2 * H, sc=0.05), R(N, K, I, H, sc=0.05)
bg, bu, Rgw, Ruw, Rgb, Rub = R(N, 2 * H), R(N, H), R(2 * H, I, sc=0.1), R(H, I, sc=0.1), R(2 * H), R(H)
mix, dY = torch.sigmoid(R(T)), R(T, N, B, H)
dims = (T, N, B, Cin, H, K)
p = lambda t: None if t is None else t.data_ptr()
st = torch.cuda.current_stream().cuda_stream
FLAGS = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for rep in range(1):
    ws = torch.zeros(lib.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4, device=dev)
    bws = torch.zeros(lib.matgcn_encoder_layer_bwd_ws_bytes(*dims, n_adp) // 4, device=dev)
    new = lambda *s: torch.zeros(*s, device=dev)
    outs = [new(T, N, B, Cin), None, new(Kp, N, ldm), new(N, K, I, 2 * H), new(N, 2 * H), new(N, K, I, H), new(N, H),
            new(2 * H, I), new(2 * H), new(H, I), new(H), new(T)]
    _cabi.check(lib.matgcn_encoder_layer_fwd(*dims, ldm, p(x), x.stride(0), None, p(M), p(Wg), p(bg), p(Wu), p(bu), p(Rgw),
                                             p(Rgb), p(Ruw), p(Rub), p(mix), p(ws), FLAGS, st), "fwd")
    _cabi.check(lib.matgcn_encoder_layer_bwd(*dims, ldm, n_adp, p(dY), dY.stride(0), p(M), p(Wg), p(Wu), p(Rgw), p(Ruw), p(mix),
                                             p(ws), p(bws), *[p(o) for o in outs], FLAGS, st), "bwd")
    torch.cuda.synchronize()
print("ok")
