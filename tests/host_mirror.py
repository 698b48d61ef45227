"""Torch (CPU) mirror of the kernel-level algorithm behind the C-ABI entry points.

TEST HELPER ONLY - the product path never imports this.  It restates, stage by
stage and buffer by buffer, what ``matgcn_encoder_layer_fwd/bwd`` and friends
compute (hoisted supports, hoisted per-node weights, time-batched input half,
reverse-time BPTT with time-batched weight gradients) so that

* the *math* of the hand-derived backward is checked on CPU against autograd
  of the oracle before any GPU minute is spent, and
* GPU tests can diff every intermediate buffer of the CUDA path against it.

Layouts are the device layouts: activations node-major ``[T, N, B, C]``,
per-node weights ``[N, K, I, O]``, base matrices ``[Kp, N, ldm]``.
"""
from __future__ import annotations

import torch

sig = torch.sigmoid


def adaptive_adj_fwd(L, Rt, ldm):
    n = L.shape[0]
    a = torch.softmax(torch.relu(L @ Rt.T), dim=1)
    out = torch.zeros(n, ldm, dtype=L.dtype)
    out[:, :n] = a
    return out


def adaptive_adj_bwd(L, Rt, A, dA):
    n = L.shape[0]
    a, da = A[:, :n], dA[:, :n]
    ds = a * (da - (da * a).sum(1, keepdim=True))
    dpre = ds * ((L @ Rt.T) > 0).to(ds.dtype)
    return dpre @ Rt, dpre.T @ L


def node_weights_fwd(E, pool, bias_pool, c):
    W = torch.einsum("nd,dkio->nkio", E, pool) * c.view(1, -1, 1, 1)
    return W, E @ bias_pool


def node_weights_bwd(E, pool, bias_pool, c, dW, db):
    G = torch.einsum("nd,nkio->dkio", E, dW)
    dpool = G * c.view(1, -1, 1, 1)
    dc = (G * pool).sum(dim=(0, 2, 3))
    dE = torch.einsum("nkio,dkio->nd", dW * c.view(1, -1, 1, 1), pool) + db @ bias_pool.T
    dbias_pool = E.T @ db
    return dE, dpool, dbias_pool, dc


def layer_fwd(x, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix):
    """x [T,N,B,Cin], h0 [N,B,H] or None, M [Kp,N,ldm] -> (y [T,N,B,H], saved)."""
    T, N, B, Cin = x.shape
    H = Rub.shape[0]
    Kp = M.shape[0]
    K = Kp + 1
    Mv = M[:, :, :N]
    dt = x.dtype
    PX = torch.zeros(T, K, N, B, Cin, dtype=dt)
    PX[:, 0] = x
    PX[:, 1:] = torch.einsum("knm,tmbc->tknbc", Mv, x)
    GX = torch.empty(T, N, B, 3 * H, dtype=dt)
    GX[..., :2 * H] = torch.einsum("tknbi,nkio->tnbo", PX, Wg[:, :, :Cin]) + bg[None, :, None, :]
    GX[..., 2 * H:] = torch.einsum("tknbi,nkio->tnbo", PX, Wu[:, :, :Cin]) + bu[None, :, None, :]
    RX = torch.empty(T, N, B, 3 * H, dtype=dt)
    RX[..., :2 * H] = x @ Rgw[:, :Cin].T + Rgb
    RX[..., 2 * H:] = x @ Ruw[:, :Cin].T + Rub
    PH = torch.zeros(T + 1, K, N, B, H, dtype=dt)
    PZ = torch.zeros(T, K, N, B, H, dtype=dt)
    if h0 is not None:
        PH[0, 0] = h0
    names = ["Z", "R", "HC", "H1", "Z2", "R2", "HC2", "ZH2"]
    sv = {k: torch.empty(T, N, B, H, dtype=dt) for k in names}
    for t in range(T):
        h = PH[t, 0]
        PH[t, 1:] = torch.einsum("knm,mbc->knbc", Mv, h)
        ag = GX[t, ..., :2 * H] + torch.einsum("knbi,nkio->nbo", PH[t], Wg[:, :, Cin:])
        z, r = sig(ag[..., :H]), sig(ag[..., H:])
        PZ[t, 0] = z * h
        PZ[t, 1:] = torch.einsum("knm,mbc->knbc", Mv, PZ[t, 0])
        au = GX[t, ..., 2 * H:] + torch.einsum("knbi,nkio->nbo", PZ[t], Wu[:, :, Cin:])
        hc = torch.tanh(au)
        h1 = r * h + (1 - r) * hc
        a2 = RX[t, ..., :2 * H] + h1 @ Rgw[:, Cin:].T
        z2, r2 = sig(a2[..., :H]), sig(a2[..., H:])
        zh2 = z2 * h1
        hc2 = torch.tanh(RX[t, ..., 2 * H:] + zh2 @ Ruw[:, Cin:].T)
        res = r2 * h1 + (1 - r2) * hc2
        PH[t + 1, 0] = mix[t] * h1 + (1 - mix[t]) * res
        for k, v in zip(names, [z, r, hc, h1, z2, r2, hc2, zh2]):
            sv[k][t] = v
    sv.update(PX=PX, GX=GX, RX=RX, PH=PH, PZ=PZ)
    return PH[1:, 0], sv


def layer_bwd(dY, sv, M, Wg, Wu, Rgw, Ruw, mix, n_adp, h0_given):
    """Reverse-time pass.  Returns dict of gradients with the C-ABI's names."""
    PX, PH, PZ = sv["PX"], sv["PH"], sv["PZ"]
    T, K, N, B, Cin = PX.shape
    H = PH.shape[-1]
    Kp = K - 1
    Mv = M[:, :, :N]
    dt = dY.dtype
    DG = torch.empty(T, N, B, 3 * H, dtype=dt)
    DR = torch.empty(T, N, B, 3 * H, dtype=dt)
    DPHa = torch.zeros(T, max(n_adp, 1), N, B, H, dtype=dt)
    DPZa = torch.zeros(T, max(n_adp, 1), N, B, H, dtype=dt)
    dmix = torch.zeros(T, dtype=dt)
    carry = torch.zeros(N, B, H, dtype=dt)
    for t in range(T - 1, -1, -1):
        h = PH[t, 0]
        z, r, hc, h1 = sv["Z"][t], sv["R"][t], sv["HC"][t], sv["H1"][t]
        z2, r2, hc2 = sv["Z2"][t], sv["R2"][t], sv["HC2"][t]
        g = mix[t]
        # B0
        dy = dY[t] + carry
        res = r2 * h1 + (1 - r2) * hc2
        dmix[t] = (dy * (h1 - res)).sum()
        dres = (1 - g) * dy
        dh1 = g * dy + dres * r2
        da3 = dres * (1 - r2) * (1 - hc2 * hc2)
        DR[t, ..., 2 * H:] = da3
        # B1
        dzh2 = da3 @ Ruw[:, Cin:]
        dh1 = dh1 + dzh2 * z2
        DR[t, ..., :H] = dzh2 * h1 * z2 * (1 - z2)
        DR[t, ..., H:2 * H] = dres * (h1 - hc2) * r2 * (1 - r2)
        # B2
        dh1 = dh1 + DR[t, ..., :2 * H] @ Rgw[:, Cin:]
        dr = dh1 * (h - hc)
        dhd = dh1 * r
        dau = dh1 * (1 - r) * (1 - hc * hc)
        DG[t, ..., 2 * H:] = dau
        DG[t, ..., H:2 * H] = dr * r * (1 - r)
        # B3
        DP = torch.einsum("nbo,nkio->knbi", dau, Wu[:, :, Cin:])
        if n_adp:
            DPZa[t] = DP[1:1 + n_adp]
        # B4
        dzh = DP[0] + torch.einsum("knm,knbc->mbc", Mv, DP[1:])
        dhd = dhd + dzh * z
        DG[t, ..., :H] = dzh * h * z * (1 - z)
        # B5
        DP = torch.einsum("nbo,nkio->knbi", DG[t, ..., :2 * H], Wg[:, :, Cin:])
        if n_adp:
            DPHa[t] = DP[1:1 + n_adp]
        # B6
        carry = dhd + DP[0] + torch.einsum("knm,knbc->mbc", Mv, DP[1:])
    out = {"dh0": carry if h0_given else None, "dmix": dmix}
    I = Cin + H
    dWg = torch.empty(N, K, I, 2 * H, dtype=dt)
    dWu = torch.empty(N, K, I, H, dtype=dt)
    dWg[:, :, Cin:] = torch.einsum("tknbi,tnbo->nkio", PH[:T], DG[..., :2 * H])
    dWu[:, :, Cin:] = torch.einsum("tknbi,tnbo->nkio", PZ, DG[..., 2 * H:])
    dWg[:, :, :Cin] = torch.einsum("tknbi,tnbo->nkio", PX, DG[..., :2 * H])
    dWu[:, :, :Cin] = torch.einsum("tknbi,tnbo->nkio", PX, DG[..., 2 * H:])
    out.update(dWg=dWg, dWu=dWu, dbg=DG[..., :2 * H].sum(dim=(0, 2)), dbu=DG[..., 2 * H:].sum(dim=(0, 2)))
    DPX = (torch.einsum("tnbo,nkio->tknbi", DG[..., :2 * H], Wg[:, :, :Cin])
           + torch.einsum("tnbo,nkio->tknbi", DG[..., 2 * H:], Wu[:, :, :Cin]))
    dX = DPX[:, 0] + torch.einsum("knm,tknbc->tmbc", Mv, DPX[:, 1:])
    dX = dX + DR[..., :2 * H] @ Rgw[:, :Cin] + DR[..., 2 * H:] @ Ruw[:, :Cin]
    out["dX"] = dX
    dM = torch.zeros_like(M)
    for a in range(n_adp):
        dM[a, :, :N] = (torch.einsum("tnbc,tmbc->nm", DPHa[:, a], PH[:T, 0])
                        + torch.einsum("tnbc,tmbc->nm", DPZa[:, a], PZ[:, 0])
                        + torch.einsum("tnbc,tmbc->nm", DPX[:, a + 1], PX[:, 0]))
    out["dM"] = dM
    dRgw = torch.empty_like(Rgw)
    dRuw = torch.empty_like(Ruw)
    dRgw[:, Cin:] = torch.einsum("tnbo,tnbi->oi", DR[..., :2 * H], sv["H1"])
    dRgw[:, :Cin] = torch.einsum("tnbo,tnbi->oi", DR[..., :2 * H], PX[:, 0])
    dRuw[:, Cin:] = torch.einsum("tnbo,tnbi->oi", DR[..., 2 * H:], sv["ZH2"])
    dRuw[:, :Cin] = torch.einsum("tnbo,tnbi->oi", DR[..., 2 * H:], PX[:, 0])
    out.update(dRgw=dRgw, dRuw=dRuw, dRgb=DR[..., :2 * H].sum(dim=(0, 1, 2)), dRub=DR[..., 2 * H:].sum(dim=(0, 1, 2)))
    out["DG"], out["DR"], out["DPX"] = DG, DR, DPX
    return out


class MirrorLayerFn(torch.autograd.Function):
    """autograd wrapper with the same signature as ``ops.EncoderLayerFn`` so tests can
    swap it in (monkeypatch) and run the host-side model logic on CPU."""

    @staticmethod
    def forward(ctx, x, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix, n_adp):
        y, sv = layer_fwd(x, h0, M, Wg, bg, Wu, bu, Rgw, Rgb, Ruw, Rub, mix)
        ctx.sv, ctx.n_adp, ctx.h0_given = sv, n_adp, h0 is not None
        ctx.save_for_backward(M, Wg, Wu, Rgw, Ruw, mix)
        return y.clone()

    @staticmethod
    def backward(ctx, dY):
        M, Wg, Wu, Rgw, Ruw, mix = ctx.saved_tensors
        g = layer_bwd(dY, ctx.sv, M, Wg, Wu, Rgw, Ruw, mix, ctx.n_adp, ctx.h0_given)
        return (g["dX"], g["dh0"], g["dM"], g["dWg"], g["dbg"], g["dWu"], g["dbu"],
                g["dRgw"], g["dRgb"], g["dRuw"], g["dRub"], g["dmix"], None)


class MirrorAdjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, L, Rt, ldm):
        A = adaptive_adj_fwd(L, Rt, ldm)
        ctx.save_for_backward(L, Rt, A)
        return A

    @staticmethod
    def backward(ctx, dA):
        L, Rt, A = ctx.saved_tensors
        dL, dRt = adaptive_adj_bwd(L, Rt, A, dA)
        return dL, dRt, None


class MirrorNodeWeightsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, E, pool, bias_pool, c):
        ctx.save_for_backward(E, pool, bias_pool, c)
        return node_weights_fwd(E, pool, bias_pool, c)

    @staticmethod
    def backward(ctx, dW, db):
        return node_weights_bwd(*ctx.saved_tensors, dW, db)


def dense_gru_layer_mirror(x, h0, Gw, Gb, Uw, Ub, flags=0):
    """gcn_off layer as plain torch autograd (node-major layout)."""
    T, N, B, Cin = x.shape
    H = Ub.shape[0]
    h = torch.zeros(N, B, H, dtype=x.dtype) if h0 is None else h0
    ys = []
    for t in range(T):
        zr = sig(torch.cat((x[t], h), -1) @ Gw.T + Gb)
        z, r = zr[..., :H], zr[..., H:]
        hc = torch.tanh(torch.cat((x[t], z * h), -1) @ Uw.T + Ub)
        h = r * h + (1 - r) * hc
        ys.append(h)
    return torch.stack(ys, 0)


def output_head_mirror(y, w, bias, p_drop=0.0, seed=0):
    """torch restatement of ops.output_head (MA.py:416-417); the counter-based mask itself is checked on the GPU
    (tests/test_gpu_train.py), here dropout is torch's."""
    y = torch.nn.functional.dropout(y, p=p_drop, training=p_drop > 0)
    tc, n, b, h = y.shape
    return torch.einsum("trh,oth->ro", y.reshape(tc, n * b, h), w) + bias[None, :]


def install(ops_module):
    """Point the three autograd entry points of ``multistgraph_b200.ops`` at the mirror
    (tests only; returns a restore callable)."""
    saved = (ops_module.encoder_layer, ops_module.adaptive_adjacency, ops_module.node_weights)
    saved_dense, saved_mm = ops_module.dense_gru_layer, ops_module.matmul
    saved_head = ops_module.output_head
    ops_module.output_head = output_head_mirror
    ops_module.dense_gru_layer = dense_gru_layer_mirror
    ops_module.matmul = lambda A, B, flags=0: A @ B
    ops_module.encoder_layer = lambda *a: MirrorLayerFn.apply(*a[:13])
    ops_module.adaptive_adjacency = lambda L, Rt, ldm: MirrorAdjFn.apply(L, Rt, ldm)
    ops_module.node_weights = lambda E, pool, bp, c, flags=0: MirrorNodeWeightsFn.apply(E, pool, bp, c)

    def restore():
        ops_module.encoder_layer, ops_module.adaptive_adjacency, ops_module.node_weights = saved
        ops_module.dense_gru_layer, ops_module.matmul = saved_dense, saved_mm
        ops_module.output_head = saved_head
    return restore
