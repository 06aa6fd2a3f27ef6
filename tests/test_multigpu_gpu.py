"""GPU (>= 2 devices): the all-reduced gradient of N batch shards equals the 1-GPU gradient of the whole batch
(SURVEY.md section 4 "Multi-GPU" tier; the reference's batch loss is mean(losses), ref GNS/main.py:284)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, S, out_path):
    import torch.distributed as dist
    import opf_graph_neural_solver_b200 as pkg
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).to(dev)
    buses, lines, gens, _ = pkg.data.make_batch(118, S, seed=21)          # same global batch on every rank
    lo, hi = pkg.parallel.shard_range(S, rank, world)
    out = model(buses[lo:hi].to(dev), lines[lo:hi].to(dev), gens[lo:hi].to(dev))
    (out[2].sum() / S).backward()                                         # shard's part of mean(total_loss)
    pkg.parallel.allreduce_gradients(model.parameters())
    sharded = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    if rank == 0:
        model.zero_grad(set_to_none=True)
        out = model(buses.to(dev), lines.to(dev), gens.to(dev))
        out[2].mean().backward()
        single = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        torch.save({"diff": float((sharded - single).abs().max()), "gmax": float(single.abs().max())}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_allreduced_shard_gradients_equal_the_single_gpu_gradient(lib, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    out_path = os.path.join(tmp_path, "parity.pt")
    mp.spawn(_rank_main, args=(world, _free_port(), 333, out_path), nprocs=world, join=True)
    r = torch.load(out_path)
    assert r["diff"] <= 1e-5 * r["gmax"], r
