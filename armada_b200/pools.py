"""One scheduling CYCLE = the pools of a cluster, one round each.

The schedule phase of a round does not shard bit-exactly — one global cost heap orders every
placement — but POOLS are independent units in the reference: `FairSchedulingAlgo.Schedule` walks
them one after the other, each with its own NodeDb, queues and job set
(scheduling/scheduling_algo.go:129-160).  With one process per GPU, rank r owns the pools
`r, r + world, r + 2·world, …` of a FIXED cycle (strong scaling: the cycle does not grow with the
number of GPUs), and the per-job claims of every pool are gathered on all ranks with one
`all_gather` (NCCL over NVLink on the GPU box, gloo in the CPU tests) — what the leader needs to
publish the cycle's result.

Within a rank the pools of a cycle run CONCURRENTLY: a round is one persistent CTA on one SM, the
library keeps up to 8 rounds of different device contexts in flight per GPU (include/armada_b200.h,
"thread safety"), so every owned pool gets its own context and its own host thread (upload → run →
download); the host side of one pool overlaps with the device rounds of the others."""
from __future__ import annotations

import threading
from typing import Callable, List, Optional, Sequence

import numpy as np


def pools_of_rank(num_pools: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership, deterministic and independent of the pool contents."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, num_pools, world))


def summary_row(p: int, res) -> List[int]:
    """[pool, scheduled, preempted, placements, checksum of the (job, node) claims]"""
    state = np.asarray(res.job_state)
    node = np.asarray(res.job_node).astype(np.int64)
    sched = state == 1
    return [p, int(res.out.num_result_scheduled), int(res.out.num_result_preempted), int(res.stats.placements),
            int((np.nonzero(sched)[0].astype(np.int64) * 1000003 + node[sched]).sum() % (2**61 - 1))]


_CLAIM_BUFFERS: dict = {}


def gather_claims(results: dict, num_pools: int, max_jobs: int, rank: int, world: int, dist=None) -> np.ndarray:
    """Every rank ends up with claims[pool][job] = node the job was scheduled on this cycle
    (0xFFFFFFFF = not scheduled): ONE all_gather of this rank's slab of per-pool claim rows.
    `results` maps pool index -> RoundResult for the pools this rank owns.  The staging buffers
    (page-locked host slabs, device slabs) are kept between cycles; with NCCL the exchange is one
    host-to-device copy, one all_gather over NVLink and one device-to-host copy."""
    per_rank = (num_pools + world - 1) // world
    single = dist is None or world == 1
    nccl = (not single) and dist.get_backend() == "nccl"
    key = (num_pools, max_jobs, rank, world, "nccl" if nccl else "host")
    buf = _CLAIM_BUFFERS.get(key)
    if buf is None:
        buf = {}
        if single:
            buf["slab"] = np.empty((per_rank, max_jobs), dtype=np.uint32)
        else:
            import torch
            buf["t_slab"] = torch.empty((per_rank, max_jobs), dtype=torch.int32, pin_memory=nccl)
            buf["slab"] = buf["t_slab"].numpy().view(np.uint32)
            buf["t_out"] = torch.empty((world, per_rank, max_jobs), dtype=torch.int32, pin_memory=nccl)
            if nccl:
                buf["d_slab"] = torch.empty((per_rank, max_jobs), dtype=torch.int32, device="cuda")
                buf["d_out"] = torch.empty((world, per_rank, max_jobs), dtype=torch.int32, device="cuda")
        _CLAIM_BUFFERS[key] = buf
    slab = buf["slab"]
    mine = pools_of_rank(num_pools, rank, world)
    for i in range(per_rank):
        row = slab[i]
        if i >= len(mine):
            row.fill(0xFFFFFFFF)
            continue
        r = results[mine[i]]
        node = np.asarray(r.job_node)
        n = len(node)
        np.copyto(row[:n], node)
        row[:n][np.asarray(r.job_state) != 1] = 0xFFFFFFFF
        row[n:].fill(0xFFFFFFFF)
    if single:
        return slab[:num_pools]
    import torch
    if nccl:
        buf["d_slab"].copy_(buf["t_slab"], non_blocking=True)
        dist.all_gather_into_tensor(buf["d_out"], buf["d_slab"])
        buf["t_out"].copy_(buf["d_out"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    else:
        parts = [torch.empty_like(buf["t_slab"]) for _ in range(world)]
        dist.all_gather(parts, buf["t_slab"])
        for w, part in enumerate(parts):
            buf["t_out"][w].copy_(part)
    out = buf["t_out"].numpy().view(np.uint32)  # [world][per_rank][max_jobs]; pool p = i * world + w sits at out[w, i]
    return out.transpose(1, 0, 2).reshape(per_rank * world, max_jobs)[:num_pools]


class PoolCycle:
    """The pools this rank owns, scheduled through `make_round()` device contexts (DeviceRound)."""

    def __init__(self, pool_inputs: Sequence, rank: int, world: int, make_round: Callable):
        self.inputs = pool_inputs
        self.rank, self.world = rank, world
        self.mine = pools_of_rank(len(pool_inputs), rank, world)
        self.make_round = make_round
        self._resident: Optional[list] = None
        self._pipe: Optional[list] = None

    # ---- inputs resident on the device: one context per owned pool ---------------------------------
    def upload_resident(self) -> None:
        if self._resident is None:
            self._resident = [self.make_round() for _ in self.mine]
        for dev, p in zip(self._resident, self.mine):
            dev.upload(self.inputs[p])

    def run_resident(self) -> list:
        """One cycle with resident inputs, the owned pools concurrently; returns the per-pool device
        statistics (every round is timed on its own stream)."""
        out: list = [None] * len(self._resident)
        err: List[BaseException] = []

        def go(i):
            try:
                out[i] = self._resident[i].run()
            except BaseException as e:
                err.append(e)

        ts = [threading.Thread(target=go, args=(i,)) for i in range(1, len(self._resident))]
        for t in ts:
            t.start()
        if self._resident:
            go(0)
        for t in ts:
            t.join()
        if err:
            raise err[0]
        return out

    def download_resident(self, i: int, res=None):
        return self._resident[i].download(res)

    # ---- host buffers in, host buffers out: one context and one host thread per owned pool ------------
    def schedule_cycle(self, results: Optional[dict] = None) -> dict:
        from .model import RoundResult
        if self._pipe is None:
            self._pipe = [self.make_round() for _ in self.mine]
        out = {} if results is None else results
        err: List[BaseException] = []

        def go(i: int, p: int):
            try:
                dev = self._pipe[i]
                res = out.get(p)
                if res is None:
                    res = RoundResult(self.inputs[p])
                dev.upload(self.inputs[p])
                res.stats = dev.run()
                dev.download(res)
                out[p] = res
            except BaseException as e:  # surfaced on the calling thread
                err.append(e)

        ts = [threading.Thread(target=go, args=(i, p)) for i, p in list(enumerate(self.mine))[1:]]
        for t in ts:
            t.start()
        if self.mine:
            go(0, self.mine[0])
        for t in ts:
            t.join()
        if err:
            raise err[0]
        return out

    def close(self):
        for d in (self._resident or []) + (self._pipe or []):
            d.close()
        self._resident = self._pipe = None


def schedule_pools(pool_inputs: Sequence, rank: int, world: int, schedule: Callable, dist=None):
    """Schedule this rank's pools with `schedule(input) -> RoundResult` and return, on every rank,
    (summary rows [pool, scheduled, preempted, placements, checksum], claims[pool][job])."""
    mine = pools_of_rank(len(pool_inputs), rank, world)
    rows = np.zeros((len(pool_inputs), 5), dtype=np.int64)
    results = {}
    for p in mine:
        results[p] = schedule(pool_inputs[p])
        rows[p] = summary_row(p, results[p])
    max_jobs = max(int(i.num_jobs) for i in pool_inputs)
    claims = gather_claims(results, len(pool_inputs), max_jobs, rank, world, dist)
    if dist is not None and world > 1:
        import torch
        t = torch.from_numpy(rows)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)  # every pool is written by exactly one rank, the others contribute zeros
        rows = t.cpu().numpy()
    return rows, claims
