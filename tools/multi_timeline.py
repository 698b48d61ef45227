"""Per-phase timeline (CTA 0) of the persistent multi-phase kernel inside encoder-layer fwd and bwd at the Baltimore shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi
lib = _cabi.lib()
dev = "cuda:0"
T, N, B, Cin, H, Kp, n_adp = 3, 403, 64, 64, 64, 4, 1
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
def run(which):
    ws = torch.zeros(lib.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4, device=dev)
    bws = torch.zeros(lib.matgcn_encoder_layer_bwd_ws_bytes(*dims, n_adp) // 4, device=dev)
    new = lambda *s: torch.zeros(*s, device=dev)
    outs = [new(T, N, B, Cin), None, new(Kp, N, ldm), new(N, K, I, 2 * H), new(N, 2 * H), new(N, K, I, H), new(N, H),
            new(2 * H, I), new(2 * H), new(H, I), new(H), new(T)]
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    if which == "fwd":
        lib.matgcn_debug_set_timeline(buf.data_ptr()); lib.matgcn_debug_set_timeline_skip(10**6)
    _cabi.check(lib.matgcn_encoder_layer_fwd(*dims, ldm, p(x), x.stride(0), None, p(M), p(Wg), p(bg), p(Wu), p(bu), p(Rgw),
                                             p(Rgb), p(Ruw), p(Rub), p(mix), p(ws), 1, st), "fwd")
    torch.cuda.synchronize()
    if which == "bwd":
        lib.matgcn_debug_set_timeline(buf.data_ptr()); lib.matgcn_debug_set_timeline_skip(10**6)
    else:
        lib.matgcn_debug_set_timeline(None)
    _cabi.check(lib.matgcn_encoder_layer_bwd(*dims, ldm, n_adp, p(dY), dY.stride(0), p(M), p(Wg), p(Wu), p(Rgw), p(Ruw), p(mix),
                                             p(ws), p(bws), *[p(o) for o in outs], 1, st), "bwd")
    torch.cuda.synchronize()
    lib.matgcn_debug_set_timeline(None)
    return buf.cpu()
run("none")
for which, names in [("fwd", ["prop", "gate", "prop", "cand", "resgate", "rescand"]),
                     ("bwd", ["head", "B1", "B2", "B3", "B4", "B5", "B6"])]:
    full = run(which)
    b = full[:1024].view(-1, 4)
    ex = full[1024:].view(-1, 16)
    t0 = b[0, 0].item()
    print("==", which, "(cycles; start = phase start after barrier wait, work = tiles done, arrive, released)")
    for i in range(len(names) * 2):
        if b[i, 0].item() == 0: break
        s_, w_, a_, r_ = [b[i, j].item() - t0 for j in range(4)]
        print(" %-8s start=%7d  work=%6d  sync_in=%5d  barrier_wait=%6d" % (names[i % len(names)], s_, w_ - s_, a_ - w_, r_ - a_))
        e = [ex[i, j].item() - t0 - s_ if ex[i, j].item() else -1 for j in range(10)]
        print("      rel. to start: producer wake=%d, issued tile0/1/2=%d/%d/%d ; mma tile0 start/commit=%d/%d tile1=%d/%d tile2=%d/%d" % tuple(e))
