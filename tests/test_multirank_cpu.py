"""CPU, world_size 2, gloo: the N>1 host logic (ensemble partition + the single all-gather)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, rows_local, ensemble_rows, q):
    sys.path.insert(0, ROOT)
    from edmp_b200 import ensemble
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    costs = torch.rand(rows_local, generator=g)
    if rank == 1:
        costs[3] = float("nan")            # a poisoned row must never be selected
        costs[ensemble_rows + 2] = -1.0
    allc = ensemble.gather_costs(costs)
    idx, best = ensemble.best_rows(allc, ensemble_rows)
    q.put((rank, allc.clone(), idx.clone(), best.clone(), list(ensemble.ensemble_slices(7, world, rank))))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_best_of_ensemble_two_ranks():
    world, rows_local, ens = 2, 12, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows_local, ens, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, all0, idx0, best0, sl0), (r1, all1, idx1, best1, sl1) = results
    assert torch.equal(torch.nan_to_num(all0, nan=-7.0), torch.nan_to_num(all1, nan=-7.0))   # same view everywhere
    assert torch.equal(idx0, idx1) and idx0.shape == (2, 2)
    g1 = torch.rand(rows_local, generator=torch.Generator().manual_seed(101))
    assert torch.equal(all0[1, :3], g1[:3]) and torch.isnan(all0[1, 3])
    assert int(idx0[1, 1]) == 2 and float(best0[1, 1]) == -1.0
    assert int(idx0[1, 0]) != 3
    ref = torch.nan_to_num(all0, nan=float("inf")).reshape(2, 2, ens).argmin(dim=2)
    assert torch.equal(idx0, ref)
    assert sl0 == [0, 1, 2, 3] and sl1 == [4, 5, 6]


def test_single_process_passthrough():
    sys.path.insert(0, ROOT)
    from edmp_b200 import ensemble
    c = torch.tensor([3.0, 1.0, float("nan"), 2.0])
    allc = ensemble.gather_costs(c)
    assert allc.shape == (1, 4)
    idx, best = ensemble.best_rows(allc, 2)
    assert idx.tolist() == [[1, 1]] and best.tolist() == [[1.0, 2.0]]
