"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (batch sharding and
the single gradient all-reduce).  The kernels themselves are exercised by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import opf_graph_neural_solver_b200 as pkg


def test_shard_range_partitions_every_batch():
    for S in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [pkg.parallel.shard_range(S, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, flat_views, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=10, hidden_dim=10, K=2, multiple_phi=True)      # parameters only; no kernels on CPU
    params = list(model.parameters())
    n = sum(p.numel() for p in params)
    gen = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(n, generator=gen)
    off = 0
    for p in params:
        g = flat[off:off + p.numel()].view(p.shape)
        p.grad = g if flat_views else g.clone()
        off += p.numel()
    if flat_views:
        assert pkg.parallel.flat_gradient(params) is not None
    pkg.parallel.allreduce_gradients(params, average=True)
    got = torch.cat([p.grad.reshape(-1) for p in params])
    want = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    out[rank] = float((got - want).abs().max())
    # sharded loss reduction: mean over the global batch == all-reduced sum of local sums / S
    S = 13
    losses = torch.arange(S, dtype=torch.float32)
    a, b = pkg.parallel.shard_range(S, rank, world)
    part = losses[a:b].sum() / S
    dist.all_reduce(part)
    assert abs(float(part) - float(losses.mean())) < 1e-6
    dist.destroy_process_group()


@pytest.mark.parametrize("flat_views", [True, False])
def test_gradient_allreduce_world2_gloo(flat_views):
    world, port = 2, _free_port()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, flat_views, out), nprocs=world, join=True)
    assert len(out) == world and max(out.values()) < 1e-6
