"""Phase timeline of the persistent forward recurrence kernel (CTA 0, first epilogue thread) at the Baltimore shape,
and CUDA-event times of the layer forward with the persistent kernel on / off.
usage: python tools/rec_timeline.py [N B Cin T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi
lib = _cabi.lib()
dev = "cuda:0"
N, B, Cin, T = (int(v) for v in (sys.argv[1:5] + ["403", "64", "64", "24"][len(sys.argv) - 1:]))
H, Kp = 64, 4
K, I = Kp + 1, Cin + H
ldm = (N + 7) // 8 * 8
g = torch.Generator().manual_seed(0)
R = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
x, M = R(T, N, B, Cin), R(Kp, N, ldm, sc=0.05)
Wg, Wu = R(N, K, I, 2 * H, sc=0.05), R(N, K, I, H, sc=0.05)
bg, bu, Rgw, Ruw, Rgb, Rub = R(N, 2 * H), R(N, H), R(2 * H, I, sc=0.1), R(H, I, sc=0.1), R(2 * H), R(H)
mix = torch.sigmoid(R(T))
dims = (T, N, B, Cin, H, K)
p = lambda t: None if t is None else t.data_ptr()
st = torch.cuda.current_stream().cuda_stream
ws = torch.zeros(lib.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4, device=dev)


def fwd():
    _cabi.check(lib.matgcn_encoder_layer_fwd(*dims, ldm, p(x), x.stride(0), None, p(M), p(Wg), p(bg), p(Wu), p(bu), p(Rgw),
                                             p(Rgb), p(Ruw), p(Rub), p(mix), p(ws), 3, st), "fwd")


for rec in (0, 1):
    lib.matgcn_set_recurrent_kernel(rec)
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fwd()
    e1.record()
    torch.cuda.synchronize()
    print("layer forward, persistent=%d: %.3f ms" % (rec, e0.elapsed_time(e1) / 5))
buf = torch.zeros(T * 16 + 512, dtype=torch.int64, device=dev)
lib.matgcn_debug_set_timeline(buf.data_ptr())
fwd()
torch.cuda.synchronize()
lib.matgcn_debug_set_timeline(None)
b = buf.cpu()[: T * 16].view(T, 4, 4)
names = ["prop h", "gate", "prop zh", "tail"]
print("cycles (CTA 0): phase start -> own tiles done | CTA barrier | grid barrier passed   [next phase start]")
for t in (0, 1, T // 2, T - 1):
    for ph in range(4):
        s, d, c, gb = (int(v) for v in b[t, ph])
        nxt = int(b[t, ph + 1, 0]) if ph < 3 else (int(b[t + 1, 0, 0]) if t + 1 < T else 0)
        print(" t=%2d %-8s tiles %6d | cta-bar %5d | grid-bar %6d | total %6d" % (t, names[ph], d - s, c - d, (gb - c) if gb else 0, (nxt - s) if nxt else 0))
tot = int(b[T - 1, 3, 0]) - int(b[0, 0, 0])
print("recurrence up to the start of the last tail: %d cycles" % tot)
f = buf.cpu()[T * 16: T * 16 + 256].view(4, 4, 16)
t0 = int(b[T // 2, 0, 0])
print("per-tile stamps at t=%d relative to the step start: producer first/last issue | mma acc-free, first operands, committed | epilogue pre-wait, acc ready, done" % (T // 2))
for ph in range(4):
    for i in range(3):
        if int(f[ph, i, 0]) == 0:
            continue
        v = [(int(x) - t0) if int(x) else -1 for x in f[ph, i]]
        print(" %-8s tile %d: prod %6d %6d | mma %6d %6d %6d | epi %6d %6d %6d" % ((names[ph], i) + tuple(v[:8]))
              + ("  [tail stages: h1 written %d, r2 ready %d, zh2 written %d, r3 ready %d]" % tuple(v[8:12]) if ph == 3 else ""))
for ph in range(4):
    d = sum(int(b[t, ph, 1]) - int(b[t, ph, 0]) for t in range(T - (ph == 3))) / (T - (ph == 3))
    w = sum(int(b[t, ph, 3]) - int(b[t, ph, 1]) for t in range(T) if int(b[t, ph, 3])) / T
    print(" mean %-8s tiles %7.0f  barriers %7.0f" % (names[ph], d, w))
