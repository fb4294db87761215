"""Pools across ranks.

The schedule phase of a round does not shard bit-exactly — one global cost heap orders every
placement — but POOLS are independent units in the reference (`FairSchedulingAlgo.Schedule` walks
them one after the other, scheduling_algo.go:129-160).  With one process per GPU, rank r schedules
the pools `r, r + world, r + 2·world, …`; there is no data-path collective, only the per-pool
results are gathered (`torch.distributed`, NCCL on the GPU box / gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def pools_of_rank(num_pools: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership, deterministic and independent of the pool contents."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, num_pools, world))


def schedule_pools(pool_inputs: Sequence, rank: int, world: int, schedule: Callable, dist=None):
    """Schedule this rank's pools with `schedule(input) -> RoundResult` and return, on every rank,
    one summary row per pool: [pool, scheduled, preempted, placements, checksum of (job, node)].

    `dist` is `torch.distributed` (already initialised) or None for a single process."""
    mine = pools_of_rank(len(pool_inputs), rank, world)
    rows = np.zeros((len(pool_inputs), 5), dtype=np.int64)
    for p in mine:
        res = schedule(pool_inputs[p])
        state = np.asarray(res.job_state)
        node = np.asarray(res.job_node).astype(np.int64)
        sched = state == 1
        rows[p] = [p, int(res.out.num_result_scheduled), int(res.out.num_result_preempted), int(res.stats.placements),
                   int((np.nonzero(sched)[0].astype(np.int64) * 1000003 + node[sched]).sum() % (2**61 - 1))]
    if dist is not None and world > 1:
        import torch
        t = torch.from_numpy(rows)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)  # every pool is written by exactly one rank, the others contribute zeros
        rows = t.cpu().numpy()
    return rows
