"""Best-of-ensemble selection across ranks (host glue around the one collective of the path).

The reference picks ``argmin`` of the per-row t=0 swept-volume cost over ONE ensemble on one device
(lib/guide.py:637-653) and has no multi-process story (README.md:24-30 only promises one).  Here
whole ensembles stay rank-local -- so the per-step whole-ensemble gradient norm (lib/guide.py:629)
never crosses ranks -- and the only exchange is an all-gather of the per-row final costs, after
which every rank knows the best row of every ensemble.

Seeding: device noise is Philox with the rank-LOCAL element index as the counter, so ranks that pass the same ``seed`` to
``Diffusion.run_steps`` / ``denoise_guided(noise="philox")`` draw identical per-step noise for equal row indices (only
x_T would differ).  Give every rank its own stream, e.g. ``seed = pass_index * world_size + rank`` (bench.py does).
"""
import torch
import torch.distributed as dist


def ensemble_slices(n_ensembles, world_size, rank):
    """Contiguous block partition of ensembles over ranks (first ranks take the remainder)."""
    base, rem = divmod(int(n_ensembles), int(world_size))
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def gather_costs(costs_local, group=None):
    """costs_local [rows_local] float32 (device tensor for NCCL, CPU tensor for gloo) ->
    [world, rows_local]; every rank must contribute the same number of rows."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return costs_local[None]
    world = dist.get_world_size(group)
    out = [torch.empty_like(costs_local) for _ in range(world)]
    dist.all_gather(out, costs_local.contiguous(), group=group)
    return torch.stack(out)


def best_rows(all_costs, ensemble_rows):
    """all_costs [world, rows_local] -> (best_row_index [world, ensembles_per_rank], best_cost); NaN
    costs (the reference's 0/0 gradient-norm poisoning) never win."""
    w, r = all_costs.shape
    c = torch.nan_to_num(all_costs, nan=float("inf")).reshape(w, r // ensemble_rows, ensemble_rows)
    best_cost, best_idx = c.min(dim=2)
    return best_idx, best_cost
