"""The fused reduction pass over DR (csrc/dr_pass.cuh: residual-cell weight and bias gradients, MA.py:142-150) against the separate
split-K contractions + column sums it replaces, through the C-ABI layer entry points, and against an fp64 evaluation of the same
sums from the workspace the backward leaves behind."""
import pytest
import torch

from multistgraph_b200 import _cabi

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _run_layer(T, N, B, Cin, H, Kp, flags, dr_pass, seed=0):
    lib = _cabi.lib()
    K, I = Kp + 1, Cin + H
    ldm = (N + 7) // 8 * 8
    g = torch.Generator().manual_seed(seed)
    R = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(DEV)   # noqa: E731
    x, M = R(T, N, B, Cin), R(Kp, N, ldm, sc=0.05)
    Wg, Wu = R(N, K, I, 2 * H, sc=0.05), R(N, K, I, H, sc=0.05)
    bg, bu, Rgw, Ruw, Rgb, Rub = R(N, 2 * H), R(N, H), R(2 * H, I, sc=0.1), R(H, I, sc=0.1), R(2 * H), R(H)
    mix, dY = torch.sigmoid(R(T)), R(T, N, B, H)
    dims = (T, N, B, Cin, H, K)
    p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
    st = torch.cuda.current_stream().cuda_stream
    ws = torch.zeros(lib.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4, device=DEV)
    bws = torch.zeros(lib.matgcn_encoder_layer_bwd_ws_bytes(*dims, 1) // 4, device=DEV)
    new = lambda *s: torch.zeros(*s, device=DEV)   # noqa: E731
    outs = dict(dx=new(T, N, B, Cin), dM=new(Kp, N, ldm), dWg=new(N, K, I, 2 * H), dbg=new(N, 2 * H), dWu=new(N, K, I, H),
                dbu=new(N, H), dRgw=new(2 * H, I), dRgb=new(2 * H), dRuw=new(H, I), dRub=new(H), dmix=new(T))
    prev = lib.matgcn_set_dr_pass(1 if dr_pass else 0)
    try:
        _cabi.check(lib.matgcn_encoder_layer_fwd(*dims, ldm, p(x), x.stride(0), None, p(M), p(Wg), p(bg), p(Wu), p(bu), p(Rgw),
                                                 p(Rgb), p(Ruw), p(Rub), p(mix), p(ws), flags, st), "fwd")
        o = outs
        _cabi.check(lib.matgcn_encoder_layer_bwd(*dims, ldm, 1, p(dY), dY.stride(0), p(M), p(Wg), p(Wu), p(Rgw), p(Ruw), p(mix),
                                                 p(ws), p(bws), p(o["dx"]), None, p(o["dM"]), p(o["dWg"]), p(o["dbg"]), p(o["dWu"]),
                                                 p(o["dbu"]), p(o["dRgw"]), p(o["dRgb"]), p(o["dRuw"]), p(o["dRub"]), p(o["dmix"]),
                                                 flags, st), "bwd")
        torch.cuda.synchronize()
    finally:
        lib.matgcn_set_dr_pass(prev)
    return outs


def _close(a, b, tol):
    scale = b.abs().max().clamp_min(1e-20)
    return float((a - b).abs().max() / scale) <= tol


@pytest.mark.parametrize("T,N,B,Cin,Kp", [
    (4, 45, 64, 64, 4),     # inner layer: hidden + input columns + bias sums in the pass
    (3, 77, 24, 64, 2),     # N*B = 1848 is not a multiple of the 32-row blocks: TMA zero-fill at the end of every time step
    (5, 45, 32, 2, 4),      # first layer (2 input channels): hidden columns only, the input side stays with its own kernel
    (2, 19, 8, 64, 1),      # fewer row blocks than SMs
])
@pytest.mark.parametrize("flags", [3, 1])   # bf16 mode and tf32 mode
def test_dr_pass_matches_separate_contractions(T, N, B, Cin, Kp, flags):
    a = _run_layer(T, N, B, Cin, 64, Kp, flags, dr_pass=True)
    b = _run_layer(T, N, B, Cin, 64, Kp, flags, dr_pass=False)
    for k in ("dRgw", "dRuw", "dRgb", "dRub"):
        # both sides are TF32 products with fp32 accumulation; only the summation order differs
        assert _close(a[k], b[k], 2e-3), k
    for k in ("dx", "dM", "dWg", "dWu", "dbg", "dbu", "dmix"):
        # (not bit-equal: several of these accumulate with atomics, in an order that differs from run to run)
        assert _close(a[k], b[k], 1e-4), k + " must not depend on the DR pass"


def test_dr_pass_launches_fewer_kernels():
    lib = _cabi.lib()
    n0 = lib.matgcn_launch_count()
    _run_layer(3, 45, 64, 64, 64, 4, 3, dr_pass=True)
    n1 = lib.matgcn_launch_count()
    _run_layer(3, 45, 64, 64, 64, 4, 3, dr_pass=False)
    n2 = lib.matgcn_launch_count()
    assert (n2 - n1) - (n1 - n0) == 4, "one pass replaces four contractions and a column sum"
