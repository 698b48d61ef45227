"""Short seeded training runs on a DC-shaped synthetic series, evaluated on a held-out test split the way the reference does:
``TrafficStateExecutor._train_epoch`` (executor:398-423: calculate_loss -> backward -> clip_grad_norm_(5) -> Adam) for the
updates, ``evaluate`` (executor:252-323: eval mode, predict, inverse scaling) and ``TrafficStateEvaluator`` in its default
'average' mode (traffic_state_evaluator.py:34-121: MAE@i / RMSE@i over the first i horizons, averaged over batches) for the
numbers.  Used by tests/test_gpu_acceptance.py and tools/train_acceptance.py to state the fast modes' downstream bound:
the north star asks for "unchanged MAE/RMSE on the DC test split"; the real split is not redistributable, so the series
comes from ``synthetic.make_series`` (SURVEY.md section 7 says so)."""
import torch

from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import StandardScaler, make_config, make_data_feature, make_series
from multistgraph_b200.train import DeviceWindowBank, FusedClipAdam, fused_train_step


def evaluator_metrics(pred_batches, true_batches):
    """MAE@i, RMSE@i as the evaluator's 'average' mode computes them (mean over batches of the per-batch value over the
    first i horizons); returns {"MAE": [..], "RMSE": [..]} of length output_window."""
    t_out = pred_batches[0].shape[1]
    mae = [0.0] * t_out
    rmse = [0.0] * t_out
    for p, y in zip(pred_batches, true_batches):
        for i in range(1, t_out + 1):
            d = p[:, :i] - y[:, :i]
            mae[i - 1] += d.abs().mean().item()
            rmse[i - 1] += d.pow(2).mean().sqrt().item()
    n = len(pred_batches)
    return {"MAE": [v / n for v in mae], "RMSE": [v / n for v in rmse]}


def train_and_evaluate(mode, n_nodes=237, batch=32, steps=150, t_out=3, hours=24 * 7 * 10, seed=0, dev="cuda:0", eval_batches=10,
                       lr=0.003, order_seed=None):
    """Returns (metrics before training, metrics after, list of training losses) for ``matgcn_mode=mode``.  Series and initial
    weights depend on ``seed``, batch order and dropout masks on ``order_seed`` (default: seed + 1), so two modes see the same
    run, and two values of ``order_seed`` give the spread that training noise alone produces.
    Learning rate: MultiStepLR as in the reference recipe (MultiATGCN.json: lr_decay, ratio applied at milestones), here
    x0.3 at 60 % and again at 80 % of the steps so that the weights settle before they are evaluated."""
    dev = torch.device(dev)
    series = make_series(n_nodes, hours, seed=seed)
    n_train = int(hours * 0.7)
    mean, std = series[:n_train, :, 0].mean().item(), series[:n_train, :, 0].std().item()   # scaler fitted on the training part
    scaled = series.clone()
    scaled[..., 0] = (scaled[..., 0] - mean) / std
    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=20, output_window=t_out, batch_size=batch, device=dev,
                      matgcn_mode=mode)
    df = make_data_feature(n_nodes, seed=seed)
    df["scaler"] = StandardScaler(mean, std)
    torch.manual_seed(seed)
    model = MultiATGCN(dict(cfg), df).to(dev)
    opt = FusedClipAdam(model.parameters(), lr=lr, eps=1e-8, max_grad_norm=5.0)
    bank = DeviceWindowBank(scaled.to(dev), 24, t_out, 2, 1, 1, 7, 28)
    starts = bank.valid_label_starts()
    train_starts = starts[starts + t_out <= n_train]
    test_starts = starts[starts >= int(hours * 0.8)]
    order_seed = seed + 1 if order_seed is None else order_seed
    g = torch.Generator().manual_seed(order_seed)
    torch.manual_seed(order_seed)   # the model draws its dropout seeds from the global generator
    gt = torch.Generator().manual_seed(seed + 2)
    perm = test_starts[torch.randperm(len(test_starts), generator=gt)]
    test_pick = [perm[i * batch:(i + 1) * batch] for i in range(min(eval_batches, len(perm) // batch))]

    def evaluate():
        model.eval()
        preds, trues = [], []
        with torch.no_grad():
            for pick in test_pick:
                b = bank.assemble(pick)
                y = model.predict(b)
                preds.append(df["scaler"].inverse_transform(y[..., :1]).float().cpu())
                trues.append(df["scaler"].inverse_transform(b["y"][..., :1]).float().cpu())
        model.train()
        return evaluator_metrics(preds, trues)

    before = evaluate()
    losses = []
    model.train()
    for it in range(steps):
        for grp in opt.param_groups:
            grp["lr"] = lr * (1.0 if it < 0.6 * steps else (0.3 if it < 0.8 * steps else 0.09))
        pick = train_starts[torch.randint(len(train_starts), (batch,), generator=g)]
        losses.append(fused_train_step(model, bank.assemble(pick), opt))
    losses = [float(v) for v in torch.stack(losses).cpu()]
    after = evaluate()
    return before, after, losses
