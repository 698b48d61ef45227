"""autograd bindings of the C-ABI operators (include/matgcn.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream and records the
backward graph; all arithmetic of these three operators runs in ``libmatgcn.so``.

* :func:`adaptive_adjacency`  - softmax(relu(L Rt^T))                     (MA.py:80-83)
* :func:`node_weights`        - per-node weights/bias from the pools       (MA.py:102-105)
* :func:`encoder_layer`       - one ATGRUEncoder layer over the window     (MA.py:200-211)

Inputs must be float32 CUDA tensors; anything else raises (no CPU fallback).
"""
from __future__ import annotations

import torch

from . import _cabi


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _cabi.MatgcnError("%s must be a CUDA tensor: the Multi-ATGCN operators have no CPU path" % name)
    if t.dtype != torch.float32:
        raise _cabi.MatgcnError("%s must be float32, got %s" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


class _AdaptiveAdjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, L, Rt, ldm):
        L, Rt = _f32c(L, "L"), _f32c(Rt, "Rt")
        n, d = L.shape
        if Rt.shape != (n, d):
            raise _cabi.MatgcnError("adaptive_adjacency: L and Rt must both be [N, D]")
        A = torch.empty(n, ldm, device=L.device, dtype=torch.float32)
        _cabi.check(_cabi.lib().matgcn_adaptive_adj_fwd(_ptr(L), _ptr(Rt), n, d, _ptr(A), ldm, _stream()),
                    "matgcn_adaptive_adj_fwd")
        ctx.save_for_backward(L, Rt, A)
        ctx.ldm = ldm
        return A

    @staticmethod
    def backward(ctx, dA):
        L, Rt, A = ctx.saved_tensors
        n, d = L.shape
        dA = _f32c(dA, "dA")
        dL = torch.empty_like(L)
        dRt = torch.empty_like(Rt)
        scratch = torch.empty(n, n, device=L.device, dtype=torch.float32)
        _cabi.check(_cabi.lib().matgcn_adaptive_adj_bwd(_ptr(L), _ptr(Rt), _ptr(A), _ptr(dA), n, d, ctx.ldm,
                                                         _ptr(dL), _ptr(dRt), _ptr(scratch), _stream()),
                    "matgcn_adaptive_adj_bwd")
        return dL, dRt, None


class _NodeWeightsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, E, pool, bias_pool, c, flags):
        E, pool, bias_pool, c = _f32c(E, "E"), _f32c(pool, "pool"), _f32c(bias_pool, "bias_pool"), _f32c(c, "c")
        ctx.flags = int(flags)
        n, d = E.shape
        d2, k, i, o = pool.shape
        if d2 != d or bias_pool.shape != (d, o) or c.shape != (k,):
            raise _cabi.MatgcnError("node_weights: inconsistent shapes")
        ctx.d = d
        if (ctx.flags & _cabi.FLAG_TF32) and d % 4:
            # embed_dim 10 (MultiATGCN.json default for the DC runs): rows of 40 bytes do not meet TMA's 16-byte pitch rule and the
            # two big products would fall back to the fp32 SIMT kernels; zero-padded to a multiple of 4 they stay on tensor cores
            dp = (d + 3) // 4 * 4
            E = torch.nn.functional.pad(E, (0, dp - d))
            pool = torch.nn.functional.pad(pool, (0, 0, 0, 0, 0, 0, 0, dp - d))
            bias_pool = torch.nn.functional.pad(bias_pool, (0, 0, 0, dp - d))
            d = dp
        W = torch.empty(n, k, i, o, device=E.device, dtype=torch.float32)
        b = torch.empty(n, o, device=E.device, dtype=torch.float32)
        _cabi.check(_cabi.lib().matgcn_nodeweights_fwd_ex(_ptr(E), _ptr(pool), _ptr(bias_pool), _ptr(c),
                                                           n, d, k, i, o, _ptr(W), _ptr(b), ctx.flags, _stream()),
                    "matgcn_nodeweights_fwd")
        ctx.save_for_backward(E, pool, bias_pool, c)
        return W, b

    @staticmethod
    def backward(ctx, dW, db):
        E, pool, bias_pool, c = ctx.saved_tensors
        n, d = E.shape
        _, k, i, o = pool.shape
        dW, db = _f32c(dW, "dW"), _f32c(db, "db")
        dE = torch.empty_like(E)
        dpool = torch.empty_like(pool)
        dbias_pool = torch.empty_like(bias_pool)
        dc = torch.empty_like(c)
        _cabi.check(_cabi.lib().matgcn_nodeweights_bwd_ex(_ptr(E), _ptr(pool), _ptr(bias_pool), _ptr(c), _ptr(dW),
                                                           _ptr(db), n, d, k, i, o, _ptr(dE), _ptr(dpool),
                                                           _ptr(dbias_pool), _ptr(dc), ctx.flags, _stream()),
                    "matgcn_nodeweights_bwd")
        if d != ctx.d:   # drop the zero padding of the embedding dimension
            dE, dpool, dbias_pool = dE[:, :ctx.d].contiguous(), dpool[:ctx.d].contiguous(), dbias_pool[:ctx.d].contiguous()
        return dE, dpool, dbias_pool, dc, None


class _EncoderLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix, n_adp, flags, prev):
        if not x.is_cuda or x.dtype != torch.float32:
            raise _cabi.MatgcnError("encoder_layer: x must be a float32 CUDA tensor (no CPU path)")
        T, N, B, Cin = x.shape
        # x may be a strided view over time (the previous layer's output inside its workspace)
        if x.stride(3) != 1 or x.stride(2) != Cin or x.stride(1) != B * Cin:
            x = x.contiguous()
        M, Wg, bg, Wu, bu = (_f32c(M, "M"), _f32c(Wg, "Wg"), _f32c(bg, "bg"), _f32c(Wu, "Wu"), _f32c(bu, "bu"))
        Rgw, Rgb, Ruw, Rub, mix = (_f32c(Rgw, "Rgw"), _f32c(Rgb, "Rgb"), _f32c(Ruw, "Ruw"), _f32c(Rub, "Rub"),
                                   _f32c(mix, "mix"))
        if h0 is not None:
            h0 = _f32c(h0, "h0")
        H = Rub.shape[0]
        Kp, n2, ldm = M.shape
        K = Kp + 1
        I = Cin + H
        if n2 != N or Wg.shape != (N, K, I, 2 * H) or Wu.shape != (N, K, I, H) or Rgw.shape != (2 * H, I) \
                or Ruw.shape != (H, I) or bg.shape != (N, 2 * H) or bu.shape != (N, H) or mix.shape != (T,):
            raise _cabi.MatgcnError("encoder_layer: inconsistent shapes")
        L = _cabi.lib()
        dims = (T, N, B, Cin, H, K)
        ws = torch.empty(L.matgcn_encoder_layer_fwd_ws_bytes(*dims) // 4, device=x.device, dtype=torch.float32)
        # layer chaining (include/matgcn.h): x is the previous layer's output view inside ITS workspace -> read it (and the bf16
        # propagated copies the previous recurrence already produced) in place
        ctx.chain = None
        if prev is not None and L.matgcn_encoder_layer_chain_ok(*dims, ldm, int(flags)):
            pws, pdims = prev
            if (pdims[0], pdims[1], pdims[2], pdims[4], pdims[5]) == (T, N, B, H, K) and pdims[4] == Cin \
                    and x.data_ptr() == pws.data_ptr() + 4 * L.matgcn_encoder_layer_y_offset(*pdims) \
                    and x.stride(0) == L.matgcn_encoder_layer_y_tstride(*pdims):
                x16 = pws.data_ptr() + 4 * L.matgcn_encoder_layer_slot_offset(b"PH16", *pdims) + 2 * K * N * B * H
                ctx.chain = (pws, x.data_ptr(), x16)     # (the tensor keeps the previous workspace alive until our backward)
        if ctx.chain is not None:
            _cabi.check(L.matgcn_encoder_layer_fwd_chained(*dims, ldm, _ptr(x), x.stride(0), ctx.chain[2], _ptr(h0), _ptr(M), _ptr(Wg),
                                                           _ptr(bg), _ptr(Wu), _ptr(bu), _ptr(Rgw), _ptr(Rgb), _ptr(Ruw), _ptr(Rub),
                                                           _ptr(mix), _ptr(ws), int(flags), _stream()), "matgcn_encoder_layer_fwd_chained")
        else:
            _cabi.check(L.matgcn_encoder_layer_fwd(*dims, ldm, _ptr(x), x.stride(0), _ptr(h0), _ptr(M), _ptr(Wg), _ptr(bg),
                                                   _ptr(Wu), _ptr(bu), _ptr(Rgw), _ptr(Rgb), _ptr(Ruw), _ptr(Rub),
                                                   _ptr(mix), _ptr(ws), int(flags), _stream()), "matgcn_encoder_layer_fwd")
        y = torch.as_strided(ws, (T, N, B, H), (L.matgcn_encoder_layer_y_tstride(*dims), B * H, H, 1),
                             L.matgcn_encoder_layer_y_offset(*dims))
        ctx.save_for_backward(M, Wg, Wu, Rgw, Ruw, mix)
        ctx.ws = ws
        ctx.dims, ctx.ldm, ctx.n_adp, ctx.has_h0, ctx.flags = dims, ldm, int(n_adp), h0 is not None, int(flags)
        _EncoderLayerFn._last_ws = (ws, dims)   # picked up by ops.encoder_layer (a Function cannot return a python tuple)
        return y

    @staticmethod
    def backward(ctx, dy):
        M, Wg, Wu, Rgw, Ruw, mix = ctx.saved_tensors
        T, N, B, Cin, H, K = ctx.dims
        I = Cin + H
        dev = dy.device
        if dy.dtype != torch.float32 or dy.stride(3) != 1 or dy.stride(2) != H or dy.stride(1) != B * H:
            dy = dy.contiguous().float()
        L = _cabi.lib()
        if ctx.ws is None:
            raise _cabi.MatgcnError("encoder_layer: backward called twice (the saved workspace was released)")
        bws = torch.empty(L.matgcn_encoder_layer_bwd_ws_bytes(*ctx.dims, ctx.n_adp) // 4, device=dev, dtype=torch.float32)
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        dx, dM = new(T, N, B, Cin), new(K - 1, N, ctx.ldm)
        dh0 = new(N, B, H) if ctx.has_h0 else None
        dWg, dbg, dWu, dbu = new(N, K, I, 2 * H), new(N, 2 * H), new(N, K, I, H), new(N, H)
        dRgw, dRgb, dRuw, dRub, dmix = new(2 * H, I), new(2 * H), new(H, I), new(H), new(T)
        if ctx.chain is not None:
            _cabi.check(L.matgcn_encoder_layer_bwd_chained(T, N, B, Cin, H, K, ctx.ldm, ctx.n_adp, _ptr(dy), dy.stride(0), _ptr(M),
                                                           _ptr(Wg), _ptr(Wu), _ptr(Rgw), _ptr(Ruw), _ptr(mix), _ptr(ctx.ws),
                                                           _ptr(bws), _ptr(dx), _ptr(dh0), _ptr(dM), _ptr(dWg), _ptr(dbg),
                                                           _ptr(dWu), _ptr(dbu), _ptr(dRgw), _ptr(dRgb), _ptr(dRuw), _ptr(dRub),
                                                           _ptr(dmix), ctx.flags, ctx.chain[1], ctx.chain[2], _stream()),
                        "matgcn_encoder_layer_bwd_chained")
        else:
            _cabi.check(L.matgcn_encoder_layer_bwd(T, N, B, Cin, H, K, ctx.ldm, ctx.n_adp, _ptr(dy), dy.stride(0), _ptr(M),
                                                   _ptr(Wg), _ptr(Wu), _ptr(Rgw), _ptr(Ruw), _ptr(mix), _ptr(ctx.ws),
                                                   _ptr(bws), _ptr(dx), _ptr(dh0), _ptr(dM), _ptr(dWg), _ptr(dbg),
                                                   _ptr(dWu), _ptr(dbu), _ptr(dRgw), _ptr(dRgb), _ptr(dRuw), _ptr(dRub),
                                                   _ptr(dmix), ctx.flags, _stream()), "matgcn_encoder_layer_bwd")
        ctx.ws = None  # GX/RX slots now hold gradients: the workspace is spent
        ctx.chain = None
        return dx, dh0, dM, dWg, dbg, dWu, dbu, dRgw, dRgb, dRuw, dRub, dmix, None, None, None


class _DenseGRULayerFn(torch.autograd.Function):
    """gcn_off ablation: plain GRU layer with node-shared nn.Linear weights (MA.py:142-150 as the main cell)."""

    @staticmethod
    def forward(ctx, x, h0, Gw, Gb, Uw, Ub, flags):
        if not x.is_cuda or x.dtype != torch.float32:
            raise _cabi.MatgcnError("dense_gru_layer: x must be a float32 CUDA tensor (no CPU path)")
        T, N, B, Cin = x.shape
        if x.stride(3) != 1 or x.stride(2) != Cin or x.stride(1) != B * Cin:
            x = x.contiguous()
        Gw, Gb, Uw, Ub = _f32c(Gw, "Gw"), _f32c(Gb, "Gb"), _f32c(Uw, "Uw"), _f32c(Ub, "Ub")
        if h0 is not None:
            h0 = _f32c(h0, "h0")
        H = Ub.shape[0]
        if Gw.shape != (2 * H, Cin + H) or Uw.shape != (H, Cin + H) or Gb.shape != (2 * H,):
            raise _cabi.MatgcnError("dense_gru_layer: inconsistent shapes")
        L = _cabi.lib()
        dims = (T, N, B, Cin, H)
        ws = torch.empty(L.matgcn_dense_gru_layer_fwd_ws_bytes(*dims) // 4, device=x.device, dtype=torch.float32)
        _cabi.check(L.matgcn_dense_gru_layer_fwd(*dims, _ptr(x), x.stride(0), _ptr(h0), _ptr(Gw), _ptr(Gb), _ptr(Uw),
                                                 _ptr(Ub), _ptr(ws), int(flags), _stream()), "matgcn_dense_gru_layer_fwd")
        y = torch.as_strided(ws, (T, N, B, H), (N * B * H, B * H, H, 1), L.matgcn_dense_gru_layer_y_offset(*dims))
        ctx.save_for_backward(x, Gw, Uw)
        ctx.ws, ctx.dims, ctx.has_h0, ctx.flags = ws, dims, h0 is not None, int(flags)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Gw, Uw = ctx.saved_tensors
        T, N, B, Cin, H = ctx.dims
        dev = dy.device
        if dy.dtype != torch.float32 or dy.stride(3) != 1 or dy.stride(2) != H or dy.stride(1) != B * H:
            dy = dy.contiguous().float()
        L = _cabi.lib()
        if ctx.ws is None:
            raise _cabi.MatgcnError("dense_gru_layer: backward called twice (the saved workspace was released)")
        bws = torch.empty(L.matgcn_dense_gru_layer_bwd_ws_bytes(*ctx.dims) // 4, device=dev, dtype=torch.float32)
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        dx = new(T, N, B, Cin)
        dh0 = new(N, B, H) if ctx.has_h0 else None
        dGw, dGb, dUw, dUb = new(2 * H, Cin + H), new(2 * H), new(H, Cin + H), new(H)
        _cabi.check(L.matgcn_dense_gru_layer_bwd(T, N, B, Cin, H, _ptr(dy), dy.stride(0), _ptr(x), x.stride(0), _ptr(Gw),
                                                 _ptr(Uw), _ptr(ctx.ws), _ptr(bws), _ptr(dx), _ptr(dh0), _ptr(dGw), _ptr(dGb),
                                                 _ptr(dUw), _ptr(dUb), ctx.flags, _stream()), "matgcn_dense_gru_layer_bwd")
        ctx.ws = None
        return dx, dh0, dGw, dGb, dUw, dUb, None


def dense_gru_layer(x, h0, Gw, Gb, Uw, Ub, flags=0):
    """gcn_off layer: x [T,N,B,Cin] node-major -> y [T,N,B,H]."""
    return _DenseGRULayerFn.apply(x, h0, Gw, Gb, Uw, Ub, flags)


class _MatmulFn(torch.autograd.Function):
    """C = A @ B for 2-D row-major float32 CUDA tensors through the library's GEMM engines (used by the
    Chebyshev recurrence T_k = 2 A T_{k-1} - T_{k-2} of the adaptive view when cheb_order > 2, MA.py:98-99)."""

    @staticmethod
    def _gemm(a_kc, b_kc, M, N, K, A, lda, Bm, ldb, flags):
        C = torch.empty(M, N, device=A.device, dtype=torch.float32)
        _cabi.check(_cabi.lib().matgcn_gemm_debug(a_kc, b_kc, M, N, K, _ptr(A), lda, _ptr(Bm), ldb, _ptr(C), N, 1,
                                                  int(flags), _stream()), "matgcn_gemm")
        return C

    @staticmethod
    def forward(ctx, A, B, flags):
        A, B = _f32c(A, "A"), _f32c(B, "B")
        M, K = A.shape
        K2, N = B.shape
        if K != K2:
            raise _cabi.MatgcnError("matmul: inner dimensions differ")
        ctx.save_for_backward(A, B)
        ctx.flags = int(flags)
        return _MatmulFn._gemm(1, 0, M, N, K, A, K, B, N, flags)

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        dC = _f32c(dC, "dC")
        M, K = A.shape
        N = B.shape[1]
        dA = _MatmulFn._gemm(1, 1, M, K, N, dC, N, B, N, ctx.flags)   # dC [M,N] x B^T : B read as [n'=k][k'=n]
        dB = _MatmulFn._gemm(0, 0, K, N, M, A, K, dC, N, ctx.flags)   # A^T [K,M] x dC [M,N]
        return dA, dB, None


def matmul(A, B, flags=0):
    """A [M,K] @ B [K,N] on the library's GEMM engines, differentiable."""
    return _MatmulFn.apply(A, B, flags)


class _OutputHeadFn(torch.autograd.Function):
    """Dropout + end_conv (MA.py:416-417, 340-344) on the node-major encoder output: matgcn_head_fwd / matgcn_head_bwd.
    The dropout mask is regenerated in the backward from the seed; nothing but (y, w) is saved."""

    @staticmethod
    def forward(ctx, y, w, bias, p_drop, seed, seed_dev):
        if not y.is_cuda or y.dtype != torch.float32:
            raise _cabi.MatgcnError("output_head: y must be a float32 CUDA tensor (no CPU path)")
        if seed_dev is not None and (not seed_dev.is_cuda or seed_dev.dtype != torch.int64 or seed_dev.numel() != 1):
            raise _cabi.MatgcnError("output_head: seed_dev must be a one-element int64 CUDA tensor")
        Tc, N, B, H = y.shape
        if y.stride(3) != 1 or y.stride(2) != H or y.stride(1) != B * H or (y.stride(0) & 3):
            y = y.contiguous()
        w, bias = _f32c(w, "w"), _f32c(bias, "bias")
        O = w.shape[0]
        if w.shape != (O, Tc, H) or bias.shape != (O,):
            raise _cabi.MatgcnError("output_head: inconsistent shapes")
        out = torch.empty(N * B, O, device=y.device, dtype=torch.float32)
        if seed_dev is None:
            _cabi.check(_cabi.lib().matgcn_head_fwd(_ptr(y), y.stride(0), Tc, N * B, H, _ptr(w), _ptr(bias), O, float(p_drop),
                                                    int(seed), _ptr(out), _stream()), "matgcn_head_fwd")
        else:
            _cabi.check(_cabi.lib().matgcn_head_fwd_dev(_ptr(y), y.stride(0), Tc, N * B, H, _ptr(w), _ptr(bias), O, float(p_drop),
                                                        int(seed), _ptr(seed_dev), _ptr(out), _stream()), "matgcn_head_fwd_dev")
        ctx.save_for_backward(y, w)
        ctx.p_drop, ctx.seed, ctx.seed_dev = float(p_drop), int(seed), seed_dev
        return out

    @staticmethod
    def backward(ctx, dout):
        y, w = ctx.saved_tensors
        Tc, N, B, H = y.shape
        O = w.shape[0]
        dout = _f32c(dout, "dout")
        dy = torch.empty(Tc, N, B, H, device=y.device, dtype=torch.float32)
        dw = torch.empty(O, Tc, H, device=y.device, dtype=torch.float32)
        db = torch.empty(O, device=y.device, dtype=torch.float32)
        if ctx.seed_dev is None:
            _cabi.check(_cabi.lib().matgcn_head_bwd(_ptr(y), y.stride(0), Tc, N * B, H, _ptr(w), O, ctx.p_drop, ctx.seed, _ptr(dout),
                                                    _ptr(dy), _ptr(dw), _ptr(db), _stream()), "matgcn_head_bwd")
        else:
            _cabi.check(_cabi.lib().matgcn_head_bwd_dev(_ptr(y), y.stride(0), Tc, N * B, H, _ptr(w), O, ctx.p_drop, ctx.seed,
                                                        _ptr(ctx.seed_dev), _ptr(dout), _ptr(dy), _ptr(dw), _ptr(db), _stream()),
                        "matgcn_head_bwd_dev")
        return dy, dw, db, None, None, None


class _MaskedMaeFn(torch.autograd.Function):
    """calculate_loss (MA.py:422-427) in two launches: StandardScaler.inverse_transform of forecast and target
    (normalization.py:62-76) + masked_mae_torch(pred, true, 0) (loss.py:17-29): matgcn_masked_mae_fwd / _bwd."""

    @staticmethod
    def forward(ctx, pred, y, mean, std, min_s):
        if not pred.is_cuda or pred.dtype != torch.float32 or not y.is_cuda or y.dtype != torch.float32:
            raise _cabi.MatgcnError("masked_mae_loss: float32 CUDA tensors only (no CPU path)")
        if pred.dim() != 4 or pred.shape != y.shape:
            raise _cabi.MatgcnError("masked_mae_loss: forecast and target must be 4-d tensors of one shape")
        import ctypes
        arr = ctypes.c_longlong * 4
        sizes, ps, ys = arr(*pred.shape), arr(*pred.stride()), arr(*y.stride())
        acc = torch.empty(3, device=pred.device, dtype=torch.float64)
        loss = torch.empty((), device=pred.device, dtype=torch.float32)
        _cabi.check(_cabi.lib().matgcn_masked_mae_fwd(_ptr(pred), _ptr(y), sizes, ps, ys, float(mean), float(std), float(min_s),
                                                      acc.data_ptr(), _ptr(loss), _stream()), "matgcn_masked_mae_fwd")
        ctx.save_for_backward(pred, y, acc)
        ctx.consts = (float(mean), float(std), float(min_s))
        return loss

    @staticmethod
    def backward(ctx, gloss):
        import ctypes
        pred, y, acc = ctx.saved_tensors
        mean, std, min_s = ctx.consts
        arr = ctypes.c_longlong * 4
        sizes, ps, ys = arr(*pred.shape), arr(*pred.stride()), arr(*y.stride())
        g = _f32c(gloss.reshape(1), "grad_loss")
        dpred = torch.empty(pred.shape, device=pred.device, dtype=torch.float32)
        _cabi.check(_cabi.lib().matgcn_masked_mae_bwd(_ptr(pred), _ptr(y), sizes, ps, ys, mean, std, min_s, acc.data_ptr(), _ptr(g),
                                                      _ptr(dpred), _stream()), "matgcn_masked_mae_bwd")
        return dpred, None, None, None, None


def masked_mae_loss(pred, y, mean, std, min_s=1e-4):
    """Scalar loss of ``calculate_loss`` for a ``StandardScaler`` with scalar statistics (forecast and target still scaled)."""
    return _MaskedMaeFn.apply(pred, y, mean, std, min_s)


HEAD_HIDDEN = 64  # rnn_units the fused head kernels are written for


def output_head(y, w, bias, p_drop=0.0, seed=0, seed_dev=None):
    """out[n*B + b, o] = bias[o] + sum_t sum_h dropout(y[t, n, b, h]) * w[o, t, h]   (y node-major [Tc, N, B, H]).
    ``seed_dev`` (one int64 on the device) is XORed into ``seed`` by the kernels: the form a captured train step uses."""
    return _OutputHeadFn.apply(y, w, bias, p_drop, seed, seed_dev)


def dropout_multipliers(n, p_drop, seed, device):
    """The 0 / scale multipliers output_head applies to elements 0..n-1 of y in [t, n, b, h] order (tests, diagnostics)."""
    m = torch.empty((n + 3) // 4 * 4, device=device, dtype=torch.float32)
    _cabi.check(_cabi.lib().matgcn_head_dropout_mask(m.numel(), float(p_drop), int(seed), _ptr(m), _stream()), "dropout_mask")
    return m[:n]


def adaptive_adjacency(L, Rt, ldm):
    """[N, ldm] row-softmax adaptive adjacency; columns >= N are zero."""
    return _AdaptiveAdjFn.apply(L, Rt, ldm)


def node_weights(E, pool, bias_pool, c, flags=0):
    """(W [N,K,I,O], b [N,O]) with the view weights c folded into W.  flags as for ``encoder_layer``."""
    return _NodeWeightsFn.apply(E, pool, bias_pool, c, flags)


def encoder_layer(x, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix, n_adp, flags=0):
    """x [T,N,B,Cin] node-major -> y [T,N,B,H] (a strided view into the layer's workspace).
    flags: _cabi.FLAG_EXACT (fp32 FFMA) or _cabi.FLAG_TF32 (tcgen05 tensor cores)."""
    y = _EncoderLayerFn.apply(x, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix, n_adp, flags, getattr(x, "_matgcn_layer", None))
    ws = getattr(_EncoderLayerFn, "_last_ws", None)
    _EncoderLayerFn._last_ws = None
    if ws is not None:
        y._matgcn_layer = ws     # (workspace tensor, dims) of the layer that produced y: lets the next layer chain onto it
    return y


MODES = {"exact": _cabi.FLAG_EXACT, "fp32": _cabi.FLAG_EXACT, "fast": _cabi.FLAG_TF32, "tf32": _cabi.FLAG_TF32,
         "bf16": _cabi.FLAG_TF32 | _cabi.FLAG_BF16}
