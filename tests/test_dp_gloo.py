"""Data-parallel path on CPU: world_size 2 over gloo.  The two-rank result (batch sharded,
one flat-bucket all-reduce) must equal the single-rank result on the concatenated batch.
The C-ABI operators are swapped for the torch mirror (tests only) because this container has
no GPU; the DP wrapper itself is device-agnostic."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multistgraph_b200 import dp, ops
from multistgraph_b200.model import MultiATGCN
from multistgraph_b200.synthetic import make_batch, make_config, make_data_feature
from tests import host_mirror

N, B, TOUT = 9, 4, 6


def _build():
    cfg = make_config(adjtype="multi", adpadj="bidirection", embed_dim=3, rnn_units=8, output_window=TOUT,
                      batch_size=B)
    df = make_data_feature(N, seed=4)
    torch.manual_seed(21)
    return MultiATGCN(cfg, df).eval(), make_batch(N, B, TOUT, seed=4)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    host_mirror.install(ops)
    model, batch = _build()
    if rank != 0:  # perturb, then check that the broadcast restores rank 0's weights
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    dp.broadcast_parameters(model)
    bucket = dp.FlatGradBucket(model.parameters())
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    shard = dp.shard_batch(batch, rank, world)
    loss = dp.train_step(model, {k: v.clone() for k, v in shard.items()}, opt, bucket, max_grad_norm=5.0)
    if rank == 0:
        torch.save({"flat": bucket.flat.clone(), "loss": loss,
                    "params": {k: v.detach().clone() for k, v in model.state_dict().items()}}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank(tmp_path):
    out = str(tmp_path / "rank0.pt")
    port = 29650 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)

    restore = host_mirror.install(ops)
    try:
        model, batch = _build()
        bucket = dp.FlatGradBucket(model.parameters())
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        dp.train_step(model, {k: v.clone() for k, v in batch.items()}, opt, bucket, max_grad_norm=5.0)
    finally:
        restore()
    assert torch.allclose(got["flat"], bucket.flat, rtol=1e-4, atol=1e-6)
    for k, v in model.state_dict().items():
        assert torch.allclose(got["params"][k], v, rtol=1e-4, atol=1e-6), k


def test_flat_bucket_views_and_zero_slots():
    restore = host_mirror.install(ops)
    try:
        model, batch = _build()
        bucket = dp.FlatGradBucket(model.parameters())
        bucket.zero()
        model.calculate_loss(batch).backward()
        off = 0
        for name, p in model.named_parameters():
            assert p.grad.data_ptr() == bucket.flat.data_ptr() + 4 * off, name
            off += p.numel()
        assert off == bucket.flat.numel()
        # bidirection: node_vec1/2 get no gradient and keep zero slots
        assert model.node_vec1.grad.abs().max().item() == 0.0
        assert bucket.flat.abs().sum().item() > 0
    finally:
        restore()


def test_shard_batch_rejects_ragged():
    with pytest.raises(ValueError):
        dp.shard_batch({"X": torch.zeros(5, 2)}, 0, 2)
