"""GPU parity of SURVEY.md 8f rows f1 (fused clip + Adam over the flat bucket) and f2 (window gather from a series
resident in HBM), through the C ABI, against the oracle (oracle/window_oracle.py) and against torch."""
import os

import numpy as np
import pytest
import torch

from oracle import window_oracle as wo

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "windows", "windows_small.npz")


def _dev():
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------------
# f2
# ------------------------------------------------------------------------------------------------
WCASES = {
    "shipped_like": (24, 24, 2, 1, 1, 7, 28, 1, 24),
    "short_out": (24, 3, 2, 1, 0, 7, 7, 1, 24),
    "closeness_only": (12, 6, 3, 0, 0, 1, 7, 1, 24),
    "half_hour": (24, 12, 1, 2, 1, 1, 3, 2, 24),
}


@pytest.mark.parametrize("name", sorted(WCASES))
def test_window_gather_bit_exact_vs_oracle_and_golden(name):
    from multistgraph_b200.train import DeviceWindowBank

    g = np.load(GOLD)
    df = g[name + "/series"]            # row sizes N*F = 10, 9, 7, 4 -> the 8-byte, scalar, scalar and 16-byte variants
    iw, ow, lc, lp, lt, ip, it, pph, hed = WCASES[name]
    x, y, starts = wo.generate_input_data(df, iw, ow, lc, lp, lt, ip, it, pph, hed)
    bank = DeviceWindowBank(torch.from_numpy(df).to(_dev()), iw, ow, lc, lp, lt, ip, it, pph, hed)
    assert torch.equal(bank.valid_label_starts(), torch.from_numpy(starts))
    out = bank.assemble(torch.from_numpy(starts), check=True)
    assert np.array_equal(out["X"].cpu().numpy(), x)
    assert np.array_equal(out["y"].cpu().numpy(), y)
    pick = g[name + "/pick"]
    assert np.array_equal(out["X"].cpu().numpy()[pick], g[name + "/x_pick"])   # the real reference's samples
    # shuffled, repeated and empty batches
    perm = torch.from_numpy(np.random.default_rng(1).permutation(np.concatenate([starts, starts[:3]])))
    out2 = bank.assemble(perm, check=True)
    idx = np.searchsorted(starts, perm.numpy())
    assert np.array_equal(out2["X"].cpu().numpy(), x[idx]) and np.array_equal(out2["y"].cpu().numpy(), y[idx])
    out0 = bank.assemble(torch.empty(0, dtype=torch.int64))
    assert out0["X"].shape[0] == 0


def test_window_gather_flags_invalid_label_starts():
    from multistgraph_b200._cabi import MatgcnError
    from multistgraph_b200.train import DeviceWindowBank

    bank = DeviceWindowBank(torch.randn(100, 3, 2, device=_dev()), 24, 24, 1, 0, 0)
    bank.assemble(torch.tensor([24, 76]), check=True)
    for bad in (23, 77, -1):
        with pytest.raises(MatgcnError):
            bank.assemble(torch.tensor([30, bad]), check=True)


def test_window_gather_full_size_checksum():
    """Baltimore shape (N=403, F=2, heads 2/1/1 at 7 d / 28 d, batch 64): every output element is a copy, so per-sample
    sums must equal sums of the corresponding series slices computed independently with torch indexing."""
    from multistgraph_b200.train import DeviceWindowBank

    torch.manual_seed(0)
    series = torch.randn(24 * 40, 403, 2, device=_dev())
    bank = DeviceWindowBank(series, 24, 24, 2, 1, 1, 7, 28)
    starts = bank.valid_label_starts()
    sel = starts[torch.randperm(len(starts))[:64]]
    out = bank.assemble(sel, check=True)
    assert out["X"].shape == (64, 96, 403, 2) and out["y"].shape == (64, 24, 403, 2)
    for b in (0, 17, 63):
        s = int(sel[b])
        want = torch.cat([series[s - o: s - o + 24] for o in bank.seg_offsets_host], 0)
        assert torch.equal(out["X"][b], want)
        assert torch.equal(out["y"][b], series[s: s + 24])


# ------------------------------------------------------------------------------------------------
# f1
# ------------------------------------------------------------------------------------------------
class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.randn(37, 5))
        self.b = torch.nn.Parameter(torch.randn(129))
        self.c = torch.nn.Parameter(torch.randn(3, 4, 7))
        self.frozen = torch.nn.Parameter(torch.randn(4), requires_grad=False)


@pytest.mark.parametrize("wd,max_norm", [(0.0, 5.0), (0.0, None), (1e-3, 0.5)])
def test_fused_clip_adam_matches_torch_and_oracle(wd, max_norm):
    """Bound: fp32 arithmetic in a different association order than torch's foreach kernels -> 2e-6 relative."""
    from multistgraph_b200.train import FusedClipAdam

    torch.manual_seed(0)
    m1 = _Toy().to(_dev())
    m2 = _Toy().to(_dev())
    m2.load_state_dict(m1.state_dict())
    ref = torch.optim.Adam([p for p in m1.parameters() if p.requires_grad], lr=0.01, eps=1e-8, weight_decay=wd)
    opt = FusedClipAdam(m2.parameters(), lr=0.01, eps=1e-8, weight_decay=wd, max_grad_norm=max_norm)
    names = [n for n, p in m2.named_parameters() if p.requires_grad]
    p64 = {n: dict(m1.named_parameters())[n].detach().double().cpu().numpy().ravel() for n in names}
    flat = np.concatenate([p64[n] for n in names])
    mo, vo = np.zeros_like(flat), np.zeros_like(flat)
    for step in range(1, 6):
        grads = {n: torch.randn_like(dict(m1.named_parameters())[n]) * (4.0 if step % 2 else 0.05) for n in names}
        opt.zero_grad()
        for n in names:
            dict(m1.named_parameters())[n].grad = grads[n].clone()
            dict(m2.named_parameters())[n].grad.add_(grads[n])       # accumulates into the flat bucket like autograd does
        if max_norm is not None:
            total = torch.nn.utils.clip_grad_norm_([p for p in m1.parameters() if p.requires_grad], max_norm)
        ref.step()
        opt.step()
        gflat = np.concatenate([grads[n].double().cpu().numpy().ravel() for n in names])
        flat, g_after, mo, vo, total_o = wo.adam_clip_reference(flat, gflat, mo, vo, step, 0.01, 0.9, 0.999, 1e-8, wd, max_norm)
        got = np.concatenate([dict(m2.named_parameters())[n].detach().cpu().numpy().ravel() for n in names])
        want = np.concatenate([dict(m1.named_parameters())[n].detach().cpu().numpy().ravel() for n in names])
        scale = np.abs(want).max()
        assert np.abs(got - want).max() / scale < 2e-6
        assert np.abs(got - flat).max() / scale < 2e-6
        if max_norm is not None:
            assert abs(float(opt.grad_norm) - float(total)) / float(total) < 1e-6
            assert abs(float(opt.grad_norm) - total_o) / total_o < 1e-6
            gm2 = np.concatenate([dict(m2.named_parameters())[n].grad.cpu().numpy().ravel() for n in names])
            assert np.abs(gm2 - g_after).max() / np.abs(g_after).max() < 1e-6       # .grad holds the clipped gradient
    assert torch.equal(m2.frozen, m1.frozen)
    # padding slots between parameters never move
    assert float(opt.param.abs().sum()) > 0 and opt.total % 64 == 0


def test_fused_train_step_matches_executor_loop_on_the_model():
    """The drop-in model trained 3 steps by fused_train_step vs the reference executor's loop body
    (zero_grad / calculate_loss / backward / clip_grad_norm_(5) / Adam.step), same seeds: forecasts stay within 1e-4."""
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature
    from multistgraph_b200.train import FusedClipAdam, fused_train_step

    dev = _dev()
    n, b = 23, 4
    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=6, output_window=6, batch_size=b, device=dev)
    df = make_data_feature(n, seed=3)
    torch.manual_seed(0)
    ma = MultiATGCN(dict(cfg), df).to(dev).eval()
    torch.manual_seed(0)
    mb = MultiATGCN(dict(cfg), df).to(dev).eval()
    mb.load_state_dict(ma.state_dict())
    oa = torch.optim.Adam(ma.parameters(), lr=0.003, eps=1e-8)
    ob = FusedClipAdam(mb.parameters(), lr=0.003, eps=1e-8, max_grad_norm=5.0)
    for i in range(3):
        batch = {k: v.to(dev) for k, v in make_batch(n, b, 6, seed=10 + i).items()}
        oa.zero_grad()
        la = ma.calculate_loss({k: v.clone() for k, v in batch.items()})
        la.backward()
        torch.nn.utils.clip_grad_norm_(ma.parameters(), 5.0)
        oa.step()
        lb = fused_train_step(mb, {k: v.clone() for k, v in batch.items()}, ob)
        assert abs(float(la) - float(lb)) / abs(float(la)) < 1e-4
    probe = {k: v.to(dev) for k, v in make_batch(n, b, 6, seed=99).items()}
    with torch.no_grad():
        ya, yb = ma.predict(probe), mb.predict(probe)
    assert (ya - yb).abs().max() / ya.abs().max() < 1e-4
    sa, sb = ma.state_dict(), mb.state_dict()
    assert list(sa) == list(sb)


# ------------------------------------------------------------------------------------------------
# f3: dropout + output head (MA.py:416-417)
# ------------------------------------------------------------------------------------------------
def _head_case(tc, n, b, o, seed=0):
    torch.manual_seed(seed)
    dev = _dev()
    h = 64
    big = torch.randn(tc, 3, n, b, h, device=dev)            # y lives strided inside a larger buffer, like the layer workspace
    y = big[:, 1]
    w = torch.randn(o, tc, h, device=dev) * 0.1
    bias = torch.randn(o, device=dev)
    return y, w, bias


@pytest.mark.parametrize("tc,n,b,o", [(24, 13, 5, 24), (24, 7, 64, 12), (1, 9, 3, 3), (5, 6, 11, 40), (24, 2, 32, 2)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_output_head_matches_torch_fp64(tc, n, b, o, p):
    """fp32 FFMA head vs an fp64 torch restatement with the SAME mask (read back through matgcn_head_dropout_mask):
    bound 2e-6 relative to the largest element (fp32 summation order only)."""
    from multistgraph_b200 import _cabi, ops

    y, w, bias = _head_case(tc, n, b, o)
    seed = 1234567
    yq, wq, bq = y.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    out = ops.output_head(yq, wq, bq, p, seed)
    assert out.shape == (n * b, o)
    g = torch.randn_like(out)
    out.backward(g)
    mult = ops.dropout_multipliers(y.numel(), p, seed, y.device).reshape(y.shape).double()
    if p == 0.0:
        assert torch.all(mult == 1.0)
    yd, wd, bd = y.double().requires_grad_(True), w.double().requires_grad_(True), bias.double().requires_grad_(True)
    ref = torch.einsum("trh,oth->ro", (yd * mult).reshape(tc, n * b, 64), wd) + bd[None, :]
    ref.backward(g.double())
    rel = lambda a, r: float((a.double() - r).abs().max() / r.abs().max().clamp_min(1e-30))  # noqa: E731
    assert rel(out, ref) < 2e-6
    assert rel(yq.grad, yd.grad) < 2e-6
    assert rel(wq.grad, wd.grad) < 2e-5        # atomically accumulated over row chunks
    assert rel(bq.grad, bd.grad) < 2e-5
    assert abs(_cabi.lib().matgcn_head_dropout_scale(0.1) - 65536.0 / (65536 - 6554)) < 1e-6


def test_dropout_mask_statistics_and_determinism():
    from multistgraph_b200 import _cabi, ops

    n = 1 << 22
    m1 = ops.dropout_multipliers(n, 0.1, 42, _dev())
    m2 = ops.dropout_multipliers(n, 0.1, 42, _dev())
    m3 = ops.dropout_multipliers(n, 0.1, 43, _dev())
    assert torch.equal(m1, m2) and not torch.equal(m1, m3)
    scale = _cabi.lib().matgcn_head_dropout_scale(0.1)
    vals = torch.unique(m1)
    assert vals.numel() == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - scale) < 1e-6
    frac = float((m1 == 0).float().mean())
    sigma = (0.1 * 0.9 / n) ** 0.5
    assert abs(frac - 6554 / 65536) < 5 * sigma
    assert abs(float(m1.mean()) - 1.0) < 5 * sigma * scale          # E[drop(x)] = x
    # no visible structure across the 8-element groups or between neighbours
    keep = (m1 != 0).float().reshape(-1, 8)
    assert float((keep.mean(0) - 0.9).abs().max()) < 6 * (0.09 / (n / 8)) ** 0.5
    c = float(((keep[:, :-1] - 0.9) * (keep[:, 1:] - 0.9)).mean())
    assert abs(c) < 6 * 0.09 / (n * 7 / 8) ** 0.5


def test_model_train_mode_is_reproducible_under_manual_seed():
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature

    dev = _dev()
    n, b = 17, 4
    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=6, output_window=6, batch_size=b, device=dev)
    df = make_data_feature(n, seed=3)
    torch.manual_seed(0)
    model = MultiATGCN(dict(cfg), df).to(dev).train()
    batch = {k: v.to(dev) for k, v in make_batch(n, b, 6, seed=1).items()}
    losses = []
    for s in (5, 5, 6):
        torch.manual_seed(s)
        losses.append(float(model.calculate_loss({k: v.clone() for k, v in batch.items()})))
    assert losses[0] == losses[1] and losses[0] != losses[2]
    model.eval()
    le = float(model.calculate_loss({k: v.clone() for k, v in batch.items()}))
    assert abs(le - losses[0]) / abs(le) < 0.2      # dropout perturbs, it does not change the scale


def test_micro_batched_step_equals_one_shot_step():
    """Gradient accumulation over batch slices (what the N = 8192 shape needs to fit) gives the same update as the one-shot
    step on labels without zeros (mask == 1): parameters agree to fp32 summation order."""
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature
    from multistgraph_b200.train import FusedClipAdam, fused_train_step

    dev = _dev()
    n, b = 19, 8
    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=6, output_window=6, batch_size=b, device=dev)
    df = make_data_feature(n, seed=4)
    models, opts = [], []
    for _ in range(2):
        torch.manual_seed(0)
        m = MultiATGCN(dict(cfg), df).to(dev).eval()
        models.append(m)
        opts.append(FusedClipAdam(m.parameters(), lr=0.003, max_grad_norm=5.0))
    batch = {k: v.to(dev) for k, v in make_batch(n, b, 6, seed=2).items()}
    # keep every label away from the |y| < 1e-4 mask threshold (loss.py:17-29): with a masked label the slices have different
    # mask densities and the slice-weighted loss legitimately differs from the global one (the fp64 oracle shows the same)
    batch["y"] = torch.where(batch["y"].abs() < 1e-2, torch.full_like(batch["y"], 0.5), batch["y"])
    l1 = fused_train_step(models[0], {k: v.clone() for k, v in batch.items()}, opts[0])
    l3 = fused_train_step(models[1], {k: v.clone() for k, v in batch.items()}, opts[1], micro_batches=3)
    assert abs(float(l1) - float(l3)) / abs(float(l1)) < 1e-5
    assert abs(float(opts[0].grad_norm) - float(opts[1].grad_norm)) / float(opts[0].grad_norm) < 1e-4
    # (the accumulated, clipped gradients are compared, not the parameters: the first Adam step moves every element by
    # lr * sign(g), which is ill-conditioned where g is ~0)
    for (k, p), (_, q) in zip(models[0].named_parameters(), models[1].named_parameters()):
        if p.grad is not None and float(p.grad.abs().max()) > 0:
            assert (p.grad - q.grad).abs().max() <= 1e-4 * p.grad.abs().max(), k


def test_fused_clip_adam_is_a_torch_optimizer_lr_schedule_and_checkpoint_interchange():
    """the executor's default recipe attaches ``MultiStepLR`` (lr_decay, steps [5,10,20,30], ratio 0.75;
    executor ``_build_lr_scheduler``) and saves / loads ``optimizer.state_dict()``.  Both must work on the fused optimiser,
    and its checkpoints must move to and from a stock ``torch.optim.Adam``."""
    from multistgraph_b200.train import FusedClipAdam

    torch.manual_seed(0)
    m1, m2 = _Toy().to(_dev()), _Toy().to(_dev())
    m2.load_state_dict(m1.state_dict())
    ref = torch.optim.Adam([p for p in m1.parameters() if p.requires_grad], lr=0.01, eps=1e-8)
    opt = FusedClipAdam(m2.parameters(), lr=0.01, eps=1e-8)
    assert isinstance(opt, torch.optim.Optimizer) and opt.param_groups[0]["lr"] == 0.01
    s1 = torch.optim.lr_scheduler.MultiStepLR(ref, milestones=[2, 4], gamma=0.75)
    s2 = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[2, 4], gamma=0.75)
    names = [n for n, p in m2.named_parameters() if p.requires_grad]

    def one_step(a, b, oa, ob, seed):
        g = torch.Generator(device="cpu").manual_seed(seed)
        oa.zero_grad()
        ob.zero_grad()
        for n in names:
            gr = torch.randn(dict(a.named_parameters())[n].shape, generator=g).to(_dev())
            dict(a.named_parameters())[n].grad = gr.clone()
            pb = dict(b.named_parameters())[n]
            if pb.grad is None:
                pb.grad = gr.clone()
            else:
                pb.grad.add_(gr)
        oa.step()
        ob.step()

    def close(a, b):
        for n in names:
            x, y = dict(a.named_parameters())[n], dict(b.named_parameters())[n]
            assert (x - y).abs().max().item() <= 2e-6 * max(1.0, y.abs().max().item()), n

    for epoch in range(5):
        one_step(m1, m2, ref, opt, epoch)
        s1.step()
        s2.step()
        assert abs(opt.param_groups[0]["lr"] - ref.param_groups[0]["lr"]) < 1e-12
    assert abs(opt.lr - 0.01 * 0.75 * 0.75) < 1e-12
    close(m1, m2)
    # fused -> torch.optim.Adam and torch.optim.Adam -> fused, then two more steps on both sides
    m3, m4 = _Toy().to(_dev()), _Toy().to(_dev())
    m3.load_state_dict(m2.state_dict())
    m4.load_state_dict(m1.state_dict())
    adam3 = torch.optim.Adam([p for p in m3.parameters() if p.requires_grad], lr=0.5)
    adam3.load_state_dict(opt.state_dict())
    fused4 = FusedClipAdam(m4.parameters(), lr=0.5)
    fused4.load_state_dict(ref.state_dict())
    assert fused4.step_count == 5 and abs(fused4.lr - ref.param_groups[0]["lr"]) < 1e-12
    for k in range(2):
        one_step(m1, m3, ref, adam3, 100 + k)
        one_step(m2, m4, opt, fused4, 100 + k)
    close(m1, m3)
    close(m1, m4)
    close(m1, m2)
    # a parameter whose storage is replaced behind the optimiser's back is an error, not a silent no-op
    from multistgraph_b200._cabi import MatgcnError

    m2.a.data = m2.a.data.clone()
    with pytest.raises(MatgcnError):
        opt.step()


@pytest.mark.parametrize("mean,std", [(0.0, 1.0), (17.1, 24.1)])
def test_fused_loss_matches_masked_mae_of_the_reference(mean, std):
    """SURVEY 8f f3, second half: inverse scaling + masked_mae_torch(pred, true, 0) (MA.py:422-427, loss.py:17-29) as one pass
    (ops.masked_mae_loss) against the torch restatement of loss.py the CPU tests pin on the reference's golden losses:
    strided forecast (the permuted view the head returns) and target (a channel slice), labels with exact zeros and with
    |value| < min_s after inverse scaling, value and gradient; and the all-masked batch, whose loss is 0 (NaN -> 0)."""
    from multistgraph_b200 import _cabi, ops
    from multistgraph_b200.model import masked_mae_torch
    from tests.util import max_rel_err

    g = torch.Generator().manual_seed(3)
    B, T, N, C = 5, 7, 33, 1
    pred_store = torch.randn(B, T, C, N, generator=g).to("cuda:0").requires_grad_(True)
    y_full = torch.randn(B, T, N, 2, generator=g)
    y_full[0, :, :5, 0] = (0.0 - mean) / std                 # inverse-scaled to exactly 0: masked
    y_full[1, 2, 7, 0] = (3e-5 - mean) / std                 # |label| < min_s after inverse scaling: masked
    y_full = y_full.to("cuda:0")
    pred = pred_store.permute(0, 1, 3, 2)                    # [B, T, N, C], not contiguous
    y = y_full[..., 0:1]
    loss = ops.masked_mae_loss(pred, y, mean, std)
    loss.backward()
    g_fused = pred_store.grad.clone()
    pred_store.grad = None
    ref = masked_mae_torch(pred * std + mean, y * std + mean, 0)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert max_rel_err(g_fused, pred_store.grad) < 2e-6
    # forward + backward are two launches of this library
    lib = _cabi.lib()
    n0 = lib.matgcn_launch_count()
    ops.masked_mae_loss(pred.detach().requires_grad_(True), y, mean, std).backward()
    assert lib.matgcn_launch_count() - n0 == 2
    # every label masked: mask / mean(mask) is NaN everywhere -> replaced by 0 -> loss 0 (loss.py:25-28)
    y0 = torch.full_like(y, (0.0 - mean) / std)
    p0 = pred.detach().clone().requires_grad_(True)
    l0 = ops.masked_mae_loss(p0, y0, mean, std)
    l0.backward()
    assert l0.item() == 0.0 and p0.grad.abs().max().item() == 0.0


# ------------------------------------------------------------------------------------------------
# f1: the whole loop body (executor:413-422) as ONE captured CUDA graph
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,n,b,slots", [("exact", 23, 4, 1), ("bf16", 40, 16, 2)])
def test_graphed_train_step_replays_match_eager_steps(mode, n, b, slots):
    """train.GraphedTrainStep: capture once, replay five times (new batch, new dropout mask, advancing Adam step count, a
    learning-rate change on the way) against five eager ``fused_train_step`` calls that are handed the same dropout keys.
    bf16 case: B = 16, H = 64 takes the persistent cooperative recurrence kernels, i.e. cooperative launches inside the graph, and
    two input slots (two graphs over one memory pool, replayed alternately: the double-buffered upload path of bench.py's e2e)."""
    from multistgraph_b200 import _cabi
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature
    from multistgraph_b200.train import FusedClipAdam, GraphedTrainStep, fused_train_step

    dev = _dev()
    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=8, output_window=6, batch_size=b, device=dev, matgcn_mode=mode)
    df = make_data_feature(n, seed=3)
    torch.manual_seed(0)
    me = MultiATGCN(dict(cfg), df).to(dev).train()
    torch.manual_seed(0)
    mg = MultiATGCN(dict(cfg), df).to(dev).train()
    mg.load_state_dict(me.state_dict())
    oe = FusedClipAdam(me.parameters(), lr=0.003, max_grad_norm=5.0)
    og = FusedClipAdam(mg.parameters(), lr=0.003, max_grad_norm=5.0)
    batches = [{k: v.to(dev) for k, v in make_batch(n, b, 6, seed=20 + i).items()} for i in range(5)]
    step = GraphedTrainStep(mg, og, batches[0], input_slots=slots)
    assert step.library_kernel_nodes > 10 and len(step.graphs) == slots
    # the warm-up steps inside the constructor must leave no trace in the training state
    for (k, p), (_, q) in zip(me.named_parameters(), mg.named_parameters()):
        assert torch.equal(p, q), k
    assert int(step._step_dev.item()) == 0 and float(og.exp_avg.abs().max()) == 0.0
    # eager arm: same key sequence (host half XOR a device half advanced by the same tick), host-side step count
    me._dropout_key_host = mg._dropout_key_host
    me._dropout_key_dev = step._key_dev.clone()
    lib = _cabi.lib()
    tol = 1e-4 if mode == "exact" else 2e-2
    losses = []
    for i, batch in enumerate(batches):
        if i == 3:
            oe.lr = og.lr = 0.0015     # what MultiStepLR does between epochs (executor:155-197)
        _cabi.check(lib.matgcn_step_tick(me._dropout_key_dev.data_ptr(), None, torch.cuda.current_stream().cuda_stream), "tick")
        le = float(fused_train_step(me, {k: v.clone() for k, v in batch.items()}, oe))
        lg = float(step({k: v.clone() for k, v in batch.items()}))
        losses.append(lg)
        assert abs(le - lg) <= tol * abs(le), (i, le, lg)
        assert abs(float(oe.grad_norm) - float(og.grad_norm)) <= 10 * tol * float(oe.grad_norm), i
    assert torch.equal(me._dropout_key_dev, step._key_dev)
    assert int(step._step_dev.item()) == 5 and og.step_count == 5 and oe.step_count == 5
    probe = {k: v.to(dev) for k, v in make_batch(n, b, 6, seed=99).items()}
    me.eval(); mg.eval()
    with torch.no_grad():
        ye, yg = me.predict(probe), mg.predict(probe)
    assert (ye - yg).abs().max() <= 20 * tol * ye.abs().max()
    # a replay with a frozen model draws a NEW dropout mask: same batch, lr = 0, different loss
    mg.train()
    og.lr = 0.0
    l1, l2 = float(step(batches[0])), float(step(batches[0]))
    assert l1 != l2 and abs(l1 - l2) < 0.2 * abs(l1)
    step.load_batch(batches[0], slot=0)
    assert float(step(slot=0)) != l2            # batch=None: the slot's buffers as they stand, again a new mask
    # checkpoints keep working: the state dict carries the host mirror of the step count; close() returns to eager stepping
    assert float(og.state_dict()["state"][0]["step"]) == 8.0
    step.close()
    og.lr = 0.003
    fused_train_step(mg, batches[1], og)
    assert og.step_count == 9


def test_graphed_train_step_refuses_what_it_cannot_capture():
    from multistgraph_b200._cabi import MatgcnError
    from multistgraph_b200.model import MultiATGCN
    from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature
    from multistgraph_b200.train import FusedClipAdam, GraphedTrainStep, fused_train_step

    dev = _dev()
    n, b = 11, 4
    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=6, output_window=6, batch_size=b, device=dev)
    model = MultiATGCN(dict(cfg), make_data_feature(n, seed=3)).to(dev).eval()
    opt = FusedClipAdam(model.parameters(), lr=0.003, max_grad_norm=5.0)
    batch = {k: v.to(dev) for k, v in make_batch(n, b, 6, seed=1).items()}
    with pytest.raises(MatgcnError):
        GraphedTrainStep(model, opt, batch)          # eval mode
    model.train()
    step = GraphedTrainStep(model, opt, batch)
    with pytest.raises(MatgcnError):
        fused_train_step(model, batch, opt)          # the step state lives on the device now
    with pytest.raises(MatgcnError):
        step({k: v[:2] for k, v in batch.items()})   # ragged batch
    step.close()
    fused_train_step(model, batch, opt)
