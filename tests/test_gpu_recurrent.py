"""Persistent recurrence kernels (csrc/rec_fwd.cuh, rec_bwd.cuh: one cooperative launch per layer and direction, grid barriers
between the four phases of a time step) against the one-launch-per-phase path of the same library.

Both paths read the same bf16 operand twins and accumulate in fp32 in the same order, so they agree far inside the
bound of the bf16 mode (tests/test_gpu_fullsize_oracle.py holds the oracle comparison at the BASELINE shapes)."""
import pytest
import torch

from multistgraph_b200 import _cabi
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature
from tests.util import clone_batch, max_rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(model, batch, lib, rec):
    prev = lib.matgcn_set_recurrent_kernel(1 if rec else 0)
    try:
        model.zero_grad(set_to_none=True)
        n0 = lib.matgcn_launch_count()
        y = model.predict(clone_batch(batch, DEV))
        n = lib.matgcn_launch_count() - n0
        model.calculate_loss(clone_batch(batch, DEV)).backward()
        torch.cuda.synchronize()
        return y.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}, n
    finally:
        lib.matgcn_set_recurrent_kernel(prev)


@pytest.mark.parametrize("N,B,adjtype,D,tout", [(70, 8, "multi", 10, 6), (37, 64, "multi", 20, 6), (21, 100, "multi", 10, 6),
                                                (150, 136, "od", 10, 3), (403, 64, "multi", 20, 24), (237, 64, "od", 10, 3),
                                                (883, 32, "multi", 20, 12)])
def test_persistent_forward_matches_per_phase_launches(N, B, adjtype, D, tout):
    """B = 8 / 64: half-height (M = 64) per-node tiles; B = 100: one full-height tile with a ragged last quadrant; B = 136:
    two row tiles per node; adjtype 'od' + bidirection: K = 2 supports; (403, 64), (237, 64): BASELINE configs 3 and 2, (883, 32):
    config 4 sharded over eight GPUs (256 / 8 samples per rank; six node tiles and four dense tiles per CTA)."""
    cfg = make_config(adjtype=adjtype, adpadj="bidirection", embed_dim=D, output_window=tout, batch_size=B,
                      device=torch.device(DEV), matgcn_mode="bf16")
    df = make_data_feature(N, seed=13)
    batch = make_batch(N, B, tout, seed=13)
    torch.manual_seed(3)
    model = MultiATGCN(dict(cfg), df).to(DEV).eval()
    lib = _cabi.lib()
    y0, g0, n0 = _run(model, batch, lib, False)
    y1, g1, n1 = _run(model, batch, lib, True)
    if B <= 64:   # (larger per-GPU batches keep the per-phase path)
        assert n0 - n1 >= 2 * (4 * 24 - 1), "the persistent kernel should replace four launches per step and layer (%d vs %d)" % (n1, n0)
    else:
        assert n0 == n1
    assert torch.isfinite(y1).all()
    # same bf16 operands and fp32 accumulation order in the streamed contractions; the residual cell's TF32 products run on
    # tcgen05 (operand truncation) here and as mma.sync (round to nearest) in the per-phase path
    errs = {"forecast": max_rel_err(y1, y0)}
    for k in g0:
        errs[k] = max_rel_err(g1[k], g0[k])
    print("[persistent vs per-phase N=%d B=%d] worst %.2e (%s), forecast %.2e" % (N, B, max(errs.values()), max(errs, key=errs.get), errs["forecast"]))
    assert errs["forecast"] < 2e-3
    bad = {k: v for k, v in errs.items() if not (v < 5e-3)}
    assert not bad, bad


@pytest.mark.parametrize("adjtype,adpadj,cheb", [("multi", "bidirection", 3), ("multi", "none", 2), ("od", "unidirection", 2)])
def test_persistent_kernels_other_support_sets(adjtype, adpadj, cheb):
    """cheb_order = 3: K = 9 supports - the forward kernel streams any K, the reverse kernel (five per-node accumulators in TMEM)
    declines and the one-launch-per-phase backward consumes what the persistent forward saved; adpadj = none: no adaptive slices
    (n_adp = 0, no per-step copies); unidirection: the adaptive view from node_vec1 / node_vec2."""
    N, B, tout = 45, 16, 6
    cfg = make_config(adjtype=adjtype, adpadj=adpadj, embed_dim=10, cheb_order=cheb, output_window=tout, batch_size=B,
                      device=torch.device(DEV), matgcn_mode="bf16")
    df = make_data_feature(N, seed=21)
    batch = make_batch(N, B, tout, seed=21)
    torch.manual_seed(4)
    model = MultiATGCN(dict(cfg), df).to(DEV).eval()
    lib = _cabi.lib()
    y0, g0, n0 = _run(model, batch, lib, False)
    y1, g1, n1 = _run(model, batch, lib, True)
    assert n0 - n1 >= 2 * (4 * 24 - 1)
    errs = {"forecast": max_rel_err(y1, y0)}
    for k in g0:
        errs[k] = max_rel_err(g1[k], g0[k])
    print("[persistent vs per-phase %s/%s cheb %d] worst %.2e (%s)" % (adjtype, adpadj, cheb, max(errs.values()), max(errs, key=errs.get)))
    assert errs["forecast"] < 2e-3
    bad = {k: v for k, v in errs.items() if not (v < 5e-3)}
    assert not bad, bad


def test_persistent_kernels_stay_inside_their_workspaces():
    """compute-sanitizer is not available on the GPU pool: the forward and reverse workspaces (everything the two persistent kernels
    write) are embedded in larger buffers whose guard bands must come back untouched - ragged shape (N not a multiple of any tile,
    batch below the 64-row tile: TMA clipping, masked rows, the last dense tile's partial rows)."""
    lib = _cabi.lib()
    T, N, B, Cin, H, Kp, n_adp = 5, 77, 24, 64, 64, 4, 1
    K, I = Kp + 1, Cin + H
    ldm = (N + 7) // 8 * 8
    g = torch.Generator().manual_seed(0)
    R = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(DEV)
    x, M = R(T, N, B, Cin), R(Kp, N, ldm, sc=0.05)
    Wg, Wu = R(N, K, I, 2 * H, sc=0.05), R(N, K, I, H, sc=0.05)
    bg, bu, Rgw, Ruw, Rgb, Rub = R(N, 2 * H), R(N, H), R(2 * H, I, sc=0.1), R(H, I, sc=0.1), R(2 * H), R(H)
    mix, dY = torch.sigmoid(R(T)), R(T, N, B, H)
    dims = (T, N, B, Cin, H, K)
    p = lambda t: None if t is None else t.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    guard, sentinel = 1 << 16, 12345.678
    nws = lib.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4
    nbws = lib.matgcn_encoder_layer_bwd_ws_bytes(*dims, n_adp) // 4
    big_ws = torch.full((nws + 2 * guard,), sentinel, device=DEV)
    big_bws = torch.full((nbws + 2 * guard,), sentinel, device=DEV)
    ws, bws = big_ws[guard:guard + nws], big_bws[guard:guard + nbws]
    ws.zero_()
    bws.zero_()
    new = lambda *s: torch.zeros(*s, device=DEV)
    outs = [new(T, N, B, Cin), None, new(Kp, N, ldm), new(N, K, I, 2 * H), new(N, 2 * H), new(N, K, I, H), new(N, H),
            new(2 * H, I), new(2 * H), new(H, I), new(H), new(T)]
    n0 = lib.matgcn_launch_count()
    _cabi.check(lib.matgcn_encoder_layer_fwd(*dims, ldm, p(x), x.stride(0), None, p(M), p(Wg), p(bg), p(Wu), p(bu), p(Rgw),
                                             p(Rgb), p(Ruw), p(Rub), p(mix), p(ws), 3, st), "fwd")
    _cabi.check(lib.matgcn_encoder_layer_bwd(*dims, ldm, n_adp, p(dY), dY.stride(0), p(M), p(Wg), p(Wu), p(Rgw), p(Ruw), p(mix),
                                             p(ws), p(bws), *[p(o) for o in outs], 3, st), "bwd")
    torch.cuda.synchronize()
    assert lib.matgcn_launch_count() - n0 < 60, "the persistent kernels did not run"
    for name, big, n in (("forward", big_ws, nws), ("reverse", big_bws, nbws)):
        assert (big[:guard] == sentinel).all() and (big[guard + n:] == sentinel).all(), "%s workspace overrun" % name
    assert torch.isfinite(ws).all() and all(torch.isfinite(o).all() for o in outs if o is not None)


@pytest.mark.parametrize("N,B,rec,mode", [(70, 16, True, "bf16"), (45, 8, False, "bf16"), (403, 64, True, "bf16"), (45, 8, False, "exact"),
                                          (70, 16, False, "tf32")])
def test_chained_inner_layer_matches_copy_and_repropagate(N, B, rec, mode):
    """Layer chaining (include/matgcn.h, matgcn_encoder_layer_fwd_chained): layer 1 reads its input and the propagated bf16
    copies M x_t where layer 0's recurrence left them (PH16[t+1, 1..K) = M h_t) instead of copying x and propagating all T
    steps again.  Same operands, same products: forecasts and every gradient agree with the unchained path (MATGCN_CHAIN=0)
    far inside the bf16 mode's bound.  rec=False: the one-launch-per-phase path produces the same twins.  The exact and tf32 modes
    chain onto the fp32 slots PH[t+1, 1..K) the same way."""
    import os

    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=10, output_window=6, batch_size=B,
                      device=torch.device(DEV), matgcn_mode=mode)
    df = make_data_feature(N, seed=5)
    batch = make_batch(N, B, 6, seed=5)
    torch.manual_seed(4)
    model = MultiATGCN(dict(cfg), df).to(DEV).eval()
    lib = _cabi.lib()
    old = os.environ.get("MATGCN_CHAIN")
    try:
        os.environ["MATGCN_CHAIN"] = "0"
        y0, g0, n0 = _run(model, batch, lib, rec)
        os.environ["MATGCN_CHAIN"] = "1"
        y1, g1, n1 = _run(model, batch, lib, rec)
    finally:
        if old is None:
            os.environ.pop("MATGCN_CHAIN", None)
        else:
            os.environ["MATGCN_CHAIN"] = old
    if mode == "bf16":
        assert n1 < n0, "the chained layer should drop the input copy (%d vs %d launches)" % (n1, n0)
    else:
        assert n1 <= n0   # (the fp32 copy was a cudaMemcpy2DAsync: not a counted launch)
    errs = {"forecast": max_rel_err(y1, y0)}
    for k in g0:
        errs[k] = max_rel_err(g1[k], g0[k])
    print("[chained vs unchained N=%d B=%d] worst %.2e (%s), forecast %.2e" % (N, B, max(errs.values()), max(errs, key=errs.get), errs["forecast"]))
    assert errs["forecast"] < 2e-3
    bad = {k: v for k, v in errs.items() if not (v < 5e-3)}
    assert not bad, bad


@pytest.mark.parametrize("N,B", [(150, 64), (300, 16)])
def test_persistent_kernels_with_column_fastest_tile_order(N, B):
    """Dense-phase tile order of the persistent kernels (rf_tile_decode): column tiles fastest is what graphs whose base-matrix stack
    exceeds L2 (N = 8192) select; forced here at shapes with several row AND column tiles, against the one-launch-per-phase path."""
    import os

    old = os.environ.get("MATGCN_REC_TN_FAST")
    os.environ["MATGCN_REC_TN_FAST"] = "1"
    try:
        test_persistent_forward_matches_per_phase_launches(N, B, "multi", 10, 6)
    finally:
        if old is None:
            os.environ.pop("MATGCN_REC_TN_FAST", None)
        else:
            os.environ["MATGCN_REC_TN_FAST"] = old
