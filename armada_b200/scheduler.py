"""Reference-shaped entry point for one scheduling round, backed ONLY by the CUDA library.

Mirrors `scheduling.NewPreemptingQueueScheduler(...).Schedule(ctx)`
(scheduling/preempting_queue_scheduler.go:50-84,84-285): the caller hands over the round's
SchedulingContext / NodeDb / job view (already flattened into an `ArmadaRoundInput`) and gets
back scheduled / preempted jobs plus the updated accounting.  There is no CPU fallback: if
libarmada_b200.so or a CUDA device is missing this raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import abi
from .model import RoundResult


class DeviceRound:
    """Owns one `ArmadaRound*` (device context).  upload → run → download, like
    populateNodeDb → Schedule → result read-back in scheduling_algo.go:740-840."""

    def __init__(self, device: int = 0, lib=None):
        # `lib` is for tests only (tests/emu_lib.py steps the same kernel source through a CPU SIMT
        # emulator); the product always loads the CUDA library and fails loudly without it.
        self.lib = lib if lib is not None else abi.load_product()
        self.h = C.c_void_p()
        self._check(self.lib.armada_round_create(device, C.byref(self.h)))
        self._input: Optional[abi.RoundInput] = None

    def _check(self, status: int):
        if status != abi.OK:
            raise abi.ArmadaError(status, f"{self.lib.armada_strerror(status).decode()}: {self.lib.armada_last_error().decode()}")

    def upload(self, inp: abi.RoundInput) -> None:
        self._check(self.lib.armada_round_upload(self.h, C.byref(inp)))
        self._input = inp

    def run(self, budget_ns: int = 0) -> abi.RoundStats:
        """`budget_ns` > 0: the cycle's maxSchedulingDuration (scheduling_algo.go:115-118); raises
        ArmadaError(E_DEADLINE) when it runs out — nothing of the round is committed."""
        stats = abi.RoundStats()
        if budget_ns:
            self._check(self.lib.armada_round_run_deadline(self.h, C.byref(stats), int(budget_ns)))
        else:
            self._check(self.lib.armada_round_run(self.h, C.byref(stats)))
        return stats

    def download(self, res: Optional[RoundResult] = None) -> RoundResult:
        if res is None:
            res = RoundResult(self._input)
        self._check(self.lib.armada_round_download(self.h, C.byref(res.out)))
        return res

    def schedule(self, inp: abi.RoundInput, res: Optional[RoundResult] = None) -> RoundResult:
        """Host buffers in, host buffers out (the end-to-end call a Go shim makes per pool)."""
        if res is None:
            res = RoundResult(inp)
        self._input = inp
        self._check(self.lib.armada_round_schedule(self.h, C.byref(inp), C.byref(res.out), C.byref(res.stats)))
        return res

    def close(self):
        if self.h:
            self.lib.armada_round_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PreemptingQueueScheduler:
    """`sch := NewPreemptingQueueScheduler(sctx, constraints, …, jobRepo, nodeDb, …); sch.Schedule(ctx)`"""

    def __init__(self, round_input: abi.RoundInput, device: int = 0):
        self.round_input = round_input
        self.device = device

    def schedule(self) -> RoundResult:
        with DeviceRound(self.device) as r:
            return r.schedule(self.round_input)


def round_schedule(inp: abi.RoundInput, device: int = 0) -> RoundResult:
    return PreemptingQueueScheduler(inp, device).schedule()


class DeviceNodeDb:
    """`NodeDb` on an empty cluster for dry-run gang checks — what SubmitChecker builds per executor
    (internal/scheduler/submitcheck.go:302-422).  `schedule_many(gangs)` = one
    `ScheduleManyWithTxn` + `Abort` per gang (gangs are lists of job-class indices), all gangs of the
    call in one kernel launch."""

    def __init__(self, inp: abi.RoundInput, device: int = 0, lib=None):
        self.lib = lib if lib is not None else abi.load_product()
        self.h = C.c_void_p()
        self._input = inp
        st = self.lib.armada_nodedb_create(device, C.byref(inp), C.byref(self.h))
        if st != abi.OK:
            raise abi.ArmadaError(st, f"{self.lib.armada_strerror(st).decode()}: {self.lib.armada_last_error().decode()}")

    def schedule_many(self, gangs):
        import numpy as np
        start = np.zeros(len(gangs) + 1, np.uint32)
        start[1:] = np.cumsum([len(g) for g in gangs])
        members = np.asarray([c for g in gangs for c in g] or [0], dtype=np.uint32)
        ok = np.zeros(max(len(gangs), 1), np.uint8)
        node = np.full(max(int(start[-1]), 1), abi.NONE, np.uint32)
        st = self.lib.armada_nodedb_schedule_many(self.h, len(gangs), start.ctypes.data_as(abi.u32p), members.ctypes.data_as(abi.u32p),
                                                  ok.ctypes.data_as(abi.u8p), node.ctypes.data_as(abi.u32p))
        if st != abi.OK:
            raise abi.ArmadaError(st, f"{self.lib.armada_strerror(st).decode()}: {self.lib.armada_last_error().decode()}")
        out_nodes = [node[int(start[g]):int(start[g + 1])].copy() for g in range(len(gangs))]
        return ok[: len(gangs)].astype(bool), out_nodes

    def select_nodes(self, classes):
        """One independent job per entry (its job class): the node it would be bound to on the empty
        cluster, abi.NONE if it fits nowhere (SelectNodeForJobWithTxn, nodedb.go:431-512)."""
        import numpy as np
        cls = np.asarray(list(classes) or [0], dtype=np.uint32)
        node = np.full(len(cls), abi.NONE, np.uint32)
        st = self.lib.armada_nodedb_select_nodes(self.h, len(list(classes)), cls.ctypes.data_as(abi.u32p), node.ctypes.data_as(abi.u32p))
        if st != abi.OK:
            raise abi.ArmadaError(st, f"{self.lib.armada_strerror(st).decode()}: {self.lib.armada_last_error().decode()}")
        return node[: len(list(classes))]

    def close(self):
        if self.h:
            self.lib.armada_nodedb_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
