"""Per-tile timelines (CTA 0) of the time-batched tensor-core launches of an encoder-layer fwd+bwd at the Baltimore shape, T = 24."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from multistgraph_b200 import _cabi
lib = _cabi.lib()
dev = "cuda:0"
T, N, B, Cin, H, Kp, n_adp = 24, 403, 64, 64, 64, 4, 1
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
names = ["prod_start", "mma_wait", "mma_start", "mma_commit", "epi_wait", "epi_start", "epi_end"]

def run(skip):
    ws = torch.zeros(lib.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4, device=dev)
    bws = torch.zeros(lib.matgcn_encoder_layer_bwd_ws_bytes(*dims, n_adp) // 4, device=dev)
    new = lambda *s: torch.zeros(*s, device=dev)
    outs = [new(T, N, B, Cin), None, new(Kp, N, ldm), new(N, K, I, 2 * H), new(N, 2 * H), new(N, K, I, H), new(N, H),
            new(2 * H, I), new(2 * H), new(H, I), new(H), new(T)]
    buf = torch.zeros(8 * 128, dtype=torch.int64, device=dev)
    if skip >= 0:
        lib.matgcn_debug_set_timeline(buf.data_ptr())
        lib.matgcn_debug_set_timeline_skip(skip)
    _cabi.check(lib.matgcn_encoder_layer_fwd(*dims, ldm, p(x), x.stride(0), None, p(M), p(Wg), p(bg), p(Wu), p(bu), p(Rgw),
                                             p(Rgb), p(Ruw), p(Rub), p(mix), p(ws), FLAGS, st), "fwd")
    _cabi.check(lib.matgcn_encoder_layer_bwd(*dims, ldm, n_adp, p(dY), dY.stride(0), p(M), p(Wg), p(Wu), p(Rgw), p(Ruw), p(mix),
                                             p(ws), p(bws), *[p(o) for o in outs], FLAGS, st), "bwd")
    torch.cuda.synchronize()
    lib.matgcn_debug_set_timeline(None)
    return buf.cpu().view(-1, 8)

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
FLAGS = int(sys.argv[2]) if len(sys.argv) > 2 else 3   # 1 = tf32, 3 = bf16
lib.matgcn_debug_set_mode(mode)
print('debug mode', mode)
run(-1)
# tensor-core launches of the default (persistent-kernel) path in order: forward PX propagation, GX, RX; backward (after the
# reverse kernel) dWg hidden, dWu hidden, dWg input, dWu input, DPX, dx dense, dx DR*R3, dM x 3
for label, skip in [("fwd PX prop", 0), ("fwd GX", 1), ("fwd RX", 2), ("bwd dWg hidden", 3), ("bwd dWu hidden", 4), ("bwd dWg input", 5),
                    ("bwd dWu input", 6), ("bwd DPX", 7), ("bwd dx dense", 8), ("bwd dx DR*R3", 9), ("bwd dM 0", 10)]:
    b = run(skip)
    t0 = b[0, 0].item()
    print("==", label)
    for i in range(10):
        if b[i, 0].item() == 0:
            break
        print(" tile %2d: " % i + " ".join("%s=%d" % (nm, b[i, j].item() - t0) for j, nm in enumerate(names)))
