"""Phase timeline of the persistent reverse-time recurrence kernel (CTA 0, first epilogue thread) at the Baltimore shape, and
CUDA-event times of the layer backward with the persistent kernels on / off.   usage: python tools/rec_bwd_timeline.py [N B Cin T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multistgraph_b200 import _cabi
lib = _cabi.lib()
dev = "cuda:0"
N, B, Cin, T = (int(v) for v in (sys.argv[1:5] + ["403", "64", "64", "24"][len(sys.argv) - 1:]))
H, Kp, n_adp = 64, 4, 1
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
ws = torch.zeros(lib.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4, device=dev)
bws = torch.zeros(lib.matgcn_encoder_layer_bwd_ws_bytes(*dims, n_adp) // 4, device=dev)
new = lambda *s: torch.zeros(*s, device=dev)
outs = [new(T, N, B, Cin), None, new(Kp, N, ldm), new(N, K, I, 2 * H), new(N, 2 * H), new(N, K, I, H), new(N, H),
        new(2 * H, I), new(2 * H), new(H, I), new(H), new(T)]


def fwd():
    _cabi.check(lib.matgcn_encoder_layer_fwd(*dims, ldm, p(x), x.stride(0), None, p(M), p(Wg), p(bg), p(Wu), p(bu), p(Rgw),
                                             p(Rgb), p(Ruw), p(Rub), p(mix), p(ws), 3, st), "fwd")


def bwd():
    _cabi.check(lib.matgcn_encoder_layer_bwd(*dims, ldm, n_adp, p(dY), dY.stride(0), p(M), p(Wg), p(Wu), p(Rgw), p(Ruw), p(mix),
                                             p(ws), p(bws), *[p(o) for o in outs], 3, st), "bwd")


for rec in (0, 1):
    lib.matgcn_set_recurrent_kernel(rec)
    tot = 0.0
    for it in range(4):
        fwd()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bwd()
        e1.record()
        torch.cuda.synchronize()
        if it:
            tot += e0.elapsed_time(e1)
    print("layer backward (recurrence + time-batched gradients), persistent=%d: %.3f ms" % (rec, tot / 3))
fwd()
torch.cuda.synchronize()
buf = torch.zeros(T * 16 + 128, dtype=torch.int64, device=dev)
lib.matgcn_debug_set_timeline(buf.data_ptr())
lib.matgcn_debug_set_timeline_skip(1 << 30)   # (keeps the per-launch timeline of gemm_tc.cuh off: only the persistent kernel writes)
bwd()
torch.cuda.synchronize()
lib.matgcn_debug_set_timeline(None)
b = buf.cpu()[: T * 16].view(T, 4, 4)
names = ["A head+B3", "B dense", "C gz+B5", "D dense"]
print("cycles (CTA 0): phase start -> own tiles done | CTA barrier | grid barrier passed")
for s in (0, 1, T // 2, T - 1):
    for ph in range(4):
        st_, d, c, gb = (int(v) for v in b[s, ph])
        print(" step %2d %-10s tiles %6d | cta-bar %5d | grid-bar %6d" % (s, names[ph], d - st_, c - d, gb - c))
print("whole recurrence: %d cycles" % (int(b[T - 1, 3, 3]) - int(b[0, 0, 0])))
for ph in range(4):
    d = sum(int(b[s, ph, 1]) - int(b[s, ph, 0]) for s in range(T)) / T
    w = sum(int(b[s, ph, 3]) - int(b[s, ph, 1]) for s in range(T)) / T
    print(" mean %-10s tiles %7.0f  barriers %7.0f" % (names[ph], d, w))
f = buf.cpu()[T * 16: T * 16 + 64].view(4, 16)
s_mid = T - 1 - T // 2
t0 = int(b[s_mid, 0, 0])
print("phase A stages at the middle step (relative to the phase start): tile start | da3 written | dzh2 ready | daz2 written | dh1 ready | gu written | DPT ready | done")
for i in range(3):
    print(" tile %d: " % i + " ".join("%6d" % (int(v) - t0) for v in f[i, :8]))
t2 = int(b[s_mid, 2, 0])
print("phase C stages: tile start | operand written | DPT ready | done")
for i in range(3):
    print(" tile %d: " % i + " ".join("%6d" % (int(v) - t2) for v in f[i, 8:12]))
