"""Fast mode (tcgen05 / TMA / TMEM, TF32 products, fp32 accumulate) against fp64 references.

Stated bound of the mode: TF32 keeps 10 mantissa bits of each operand, so a single contraction
is accurate to ~1e-3 relative to the output scale; through 48 recurrent steps and the backward
pass the bound asserted here is 1e-2 (max|a-b| / max|b| per tensor)."""
import numpy as np
import pytest
import torch

from multistgraph_b200 import ops  # noqa: E402
from multistgraph_b200 import _cabi
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature
from oracle.matgcn_oracle import OracleModel
from tests.util import clone_batch, max_rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GEMM_TOL = 3e-3
MODEL_TOL = 1e-2


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).float()


@pytest.mark.parametrize("a_kc,b_kc", [(1, 0), (0, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,splits", [(128, 128, 32, 1), (128, 128, 256, 1), (403, 4096, 403, 1), (64, 64, 320, 1),
                                          (1612, 520, 403, 1), (37, 72, 100, 1), (403, 403, 4096, 4), (300, 64, 2000, 7)])
def test_gemm_engine_tf32(a_kc, b_kc, M, N, K, splits):
    """Every operand-layout combination the path uses, ragged sizes, split-K - and the call must
    really have gone to the tensor-core kernel (no silent SIMT fallback)."""
    lib = _cabi.lib()
    pad = lambda v: (v + 3) // 4 * 4  # noqa: E731  (TMA needs 16-byte row pitches)
    A = _rand(M, K, seed=1)
    B = _rand(K, N, seed=2)
    ref = (A.double() @ B.double())
    if a_kc:
        Ad = torch.zeros(M, pad(K)); Ad[:, :K] = A; lda = pad(K)
    else:
        Ad = torch.zeros(K, pad(M)); Ad[:, :M] = A.t(); lda = pad(M)
    if b_kc:
        Bd = torch.zeros(N, pad(K)); Bd[:, :K] = B.t(); ldb = pad(K)
    else:
        Bd = torch.zeros(K, pad(N)); Bd[:, :N] = B; ldb = pad(N)
    Ad, Bd = Ad.to(DEV), Bd.to(DEV)
    ldc = N + 3
    C = torch.full((M, ldc), float("nan"), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    before = lib.matgcn_tc_launch_count()
    _cabi.check(lib.matgcn_gemm_debug(a_kc, b_kc, M, N, K, Ad.data_ptr(), lda, Bd.data_ptr(), ldb, C.data_ptr(), ldc,
                                      splits, _cabi.FLAG_TF32, st), "gemm_debug")
    torch.cuda.synchronize()
    assert lib.matgcn_tc_launch_count() == before + 1, "fell back to the SIMT engine"
    err = max_rel_err(C[:, :N], ref)
    print("[tc gemm a_kc=%d b_kc=%d %dx%dx%d s=%d] err=%.2e" % (a_kc, b_kc, M, N, K, splits, err))
    assert err < GEMM_TOL
    assert torch.isnan(C[:, N:]).all(), "wrote outside the tile bounds"
    # same call through the exact engine for reference
    C2 = torch.empty(M, ldc, device=DEV)
    _cabi.check(lib.matgcn_gemm_debug(a_kc, b_kc, M, N, K, Ad.data_ptr(), lda, Bd.data_ptr(), ldb, C2.data_ptr(), ldc,
                                      splits, _cabi.FLAG_EXACT, st), "gemm_debug")
    assert max_rel_err(C2[:, :N], ref) < 1e-5


BF16_TOL = 3e-2  # bf16 keeps 8 mantissa bits of the propagated operands


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("N,B,adjtype,adpadj,D,tout", [(45, 8, "multi", "bidirection", 20, 24),
                                                        (130, 4, "od", "bidirection", 10, 3)])
def test_model_fast_mode_matches_oracle(N, B, adjtype, adpadj, D, tout, mode):
    cfg = make_config(adjtype=adjtype, adpadj=adpadj, embed_dim=D, output_window=tout, batch_size=B,
                      device=torch.device(DEV), matgcn_mode=mode)
    df = make_data_feature(N, seed=5)
    batch = make_batch(N, B, tout, seed=5)
    torch.manual_seed(0)
    model = MultiATGCN(dict(cfg), df).to(DEV).eval()
    ora = OracleModel(cfg, df, {k: v.cpu() for k, v in model.state_dict().items()}, dtype=torch.float64)
    y_ref = ora.forward(clone_batch(batch))
    loss_ref = ora.calculate_loss(clone_batch(batch))
    loss_ref.backward()
    lib = _cabi.lib()
    before = lib.matgcn_tc_launch_count()
    y = model.predict(clone_batch(batch, DEV))
    loss = model.calculate_loss(clone_batch(batch, DEV))
    loss.backward()
    torch.cuda.synchronize()
    assert lib.matgcn_tc_launch_count() > before + 20, "fast mode did not use the tensor-core kernels"
    errs = {"forecast": max_rel_err(y, y_ref), "loss": abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())}
    grads = ora.grads()
    for k, p in model.named_parameters():
        if grads.get(k) is not None and p.grad is not None:
            errs["d" + k] = max_rel_err(p.grad, grads[k])
    print("[fast %s N=%d] " % (mode, N) + ", ".join("%s=%.2e" % kv for kv in errs.items()))
    tol = MODEL_TOL if mode == "tf32" else BF16_TOL
    bad = {k: v for k, v in errs.items() if not (v < tol)}
    assert not bad, bad


@pytest.mark.parametrize("N,D,K,I,O", [(33, 10, 5, 66, 128), (50, 20, 5, 128, 64), (403, 20, 5, 128, 128)])
def test_node_weights_op_tensor_core(N, D, K, I, O):
    """Per-node weight generation and its backward with the big products on the tensor-core engine (TF32), against
    the torch mirror of the operator."""
    from tests import host_mirror as hm
    E, pool, bp = _rand(N, D, seed=1), _rand(D, K, I, O, seed=2) * 0.1, _rand(D, O, seed=3)
    c = torch.softmax(_rand(K, seed=4), 0)
    dW, db = _rand(N, K, I, O, seed=5), _rand(N, O, seed=6)
    ref_in = [t.clone().requires_grad_(True) for t in (E, pool, bp, c)]
    W_ref, b_ref = hm.MirrorNodeWeightsFn.apply(*ref_in)
    torch.autograd.backward([W_ref, b_ref], [dW, db])
    gpu_in = [t.to(DEV).requires_grad_(True) for t in (E, pool, bp, c)]
    n0 = _cabi.lib().matgcn_tc_launch_count()
    W, b = ops.node_weights(*gpu_in, _cabi.FLAG_TF32)
    torch.autograd.backward([W, b], [dW.to(DEV), db.to(DEV)])
    torch.cuda.synchronize()
    # (TMA needs 16-byte row pitches: D = 10, MultiATGCN.json's embed_dim, is zero-padded to 12 inside the operator)
    assert _cabi.lib().matgcn_tc_launch_count() - n0 >= 3, "the big products did not run on the tensor-core engine"
    assert all(a.grad.shape == r.grad.shape for a, r in zip(gpu_in, ref_in))
    errs = {"W": max_rel_err(W, W_ref), "b": max_rel_err(b, b_ref)}
    for nm, a, r in zip(["dE", "dpool", "dbias_pool", "dc"], gpu_in, ref_in):
        errs[nm] = max_rel_err(a.grad, r.grad)
    print("[tc nodeweights N=%d] " % N + ", ".join("%s=%.2e" % kv for kv in errs.items()))
    bad = {k: v for k, v in errs.items() if not (v < 2e-3)}
    assert not bad, bad


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("N,B", [(70, 8), (37, 64), (21, 100)])
def test_fused_step_tail_matches_separate_contractions(N, B, mode):
    """The fused tail of the forward step (candidate contraction + residual GRU cell + mix in one launch, the residual
    products as mma.sync inside the epilogue warps) against the three separate tensor-core contractions: same
    forecasts and gradients up to TF32 rounding of the residual products (RNA in mma.sync vs truncation in tcgen05).
    B = 8 / 64 exercise the half-height (M = 64) tiles, B = 100 the full-height ones with a ragged last quadrant."""
    tout = 6
    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=10, output_window=tout, batch_size=B,
                      device=torch.device(DEV), matgcn_mode=mode)
    df = make_data_feature(N, seed=11)
    batch = make_batch(N, B, tout, seed=11)
    torch.manual_seed(2)
    model = MultiATGCN(dict(cfg), df).to(DEV).eval()
    lib = _cabi.lib()

    def run(fused):
        prev = lib.matgcn_set_fused_tail(1 if fused else 0)
        prev_rec = lib.matgcn_set_recurrent_kernel(0)   # compare the per-phase launches (the persistent kernel always fuses the tail)
        try:
            model.zero_grad(set_to_none=True)
            n0 = lib.matgcn_launch_count()
            y = model.predict(clone_batch(batch, DEV))
            n = lib.matgcn_launch_count() - n0
            model.calculate_loss(clone_batch(batch, DEV)).backward()
            torch.cuda.synchronize()
            return y.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}, n
        finally:
            lib.matgcn_set_fused_tail(prev)
            lib.matgcn_set_recurrent_kernel(prev_rec)

    y0, g0, n0 = run(False)
    y1, g1, n1 = run(True)
    assert n0 - n1 == 2 * 24 * 2, "fused tail should save two launches per step and layer (%d vs %d)" % (n1, n0)
    assert max_rel_err(y1, y0) < 2e-3
    for k in g0:
        assert max_rel_err(g1[k], g0[k]) < MODEL_TOL, k


@pytest.mark.parametrize("a_kc,b_kc", [(1, 0), (0, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,splits", [(128, 128, 64, 1), (403, 4096, 403, 1), (64, 64, 320, 1), (1612, 520, 403, 1),
                                          (37, 72, 100, 1), (403, 403, 4096, 4)])
def test_gemm_engine_bf16(a_kc, b_kc, M, N, K, splits):
    """bf16 operand twins on the bf16 tcgen05 engine (kind::f16), fp32 accumulation; compared with the fp64 product
    of the SAME bf16-rounded operands, so only accumulation order differs."""
    lib = _cabi.lib()
    pad = lambda v: (v + 7) // 8 * 8  # noqa: E731  (16-byte row pitches in bf16)
    A = _rand(M, K, seed=1).bfloat16()
    B = _rand(K, N, seed=2).bfloat16()
    ref = A.double() @ B.double()
    if a_kc:
        Ad = torch.zeros(M, pad(K), dtype=torch.bfloat16); Ad[:, :K] = A; lda = pad(K)
    else:
        Ad = torch.zeros(K, pad(M), dtype=torch.bfloat16); Ad[:, :M] = A.t(); lda = pad(M)
    if b_kc:
        Bd = torch.zeros(N, pad(K), dtype=torch.bfloat16); Bd[:, :K] = B.t(); ldb = pad(K)
    else:
        Bd = torch.zeros(K, pad(N), dtype=torch.bfloat16); Bd[:, :N] = B; ldb = pad(N)
    Ad, Bd = Ad.to(DEV), Bd.to(DEV)
    ldc = N + 3
    C = torch.full((M, ldc), float("nan"), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.matgcn_gemm_debug_bf16(a_kc, b_kc, M, N, K, Ad.data_ptr(), lda, Bd.data_ptr(), ldb, C.data_ptr(), ldc,
                                           splits, st), "gemm_debug_bf16")
    torch.cuda.synchronize()
    err = max_rel_err(C[:, :N], ref)
    print("[bf16 gemm a_kc=%d b_kc=%d %dx%dx%d s=%d] err=%.2e" % (a_kc, b_kc, M, N, K, splits, err))
    assert err < 1e-5
    assert torch.isnan(C[:, N:]).all()


@pytest.mark.parametrize("mode,tol", [("tf32", MODEL_TOL), ("bf16", BF16_TOL)])
def test_full_size_fast_modes_match_exact_mode(mode, tol):
    """Baltimore shape (N=403, K=5, D=20, 24 -> 24) at batch 16: the fast modes against the exact fp32 engine of the same
    library (itself pinned to the oracle at 1e-4) - the only place where the full-size tile shapes, the bf16-only slots, the
    split epilogue and the mma.sync input-side gradients are all exercised together.  Bounds: those of the mode on every
    parameter gradient, 2e-3 on the forecast."""
    from multistgraph_b200.synthetic import workload

    cfg, df, batch = workload("baltimore_multi", seed=0, batch=16)
    outs = {}
    for m in ("exact", mode):
        c = dict(cfg)
        c["matgcn_mode"] = m
        c["device"] = torch.device(DEV)
        torch.manual_seed(0)
        model = MultiATGCN(c, df).to(DEV).eval()
        y = model.predict(clone_batch(batch, DEV))
        model.calculate_loss(clone_batch(batch, DEV)).backward()
        torch.cuda.synchronize()
        outs[m] = (y.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})
    y0, g0 = outs["exact"]
    y1, g1 = outs[mode]
    errs = {"forecast": max_rel_err(y1, y0)}
    for k in g0:
        errs["d" + k] = max_rel_err(g1[k], g0[k])
    print("[full size %s vs exact] " % mode + ", ".join("%s=%.2e" % kv for kv in sorted(errs.items(), key=lambda kv: -kv[1])[:6]))
    assert errs["forecast"] < 2e-3
    bad = {k: v for k, v in errs.items() if not (v < tol)}
    assert not bad, bad
