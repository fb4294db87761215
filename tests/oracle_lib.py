"""Loader for the CPU oracle (oracle/libarmada_oracle.so).  TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from armada_b200 import abi
from armada_b200.model import RoundResult

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libarmada_oracle.so")
_lib = None


def build(force: bool = False) -> None:
    src = os.path.join(ORACLE_DIR, "armada_oracle.cpp")
    stale = (not os.path.exists(ORACLE_LIB)) or any(
        os.path.getmtime(p) > os.path.getmtime(ORACLE_LIB)
        for p in (src, os.path.join(ORACLE_DIR, "armada_oracle.h"), os.path.join(ROOT, "include", "armada_b200.h"))
        if os.path.exists(p)
    )
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if os.path.exists(os.path.join(ORACLE_DIR, "armada_oracle.cpp")):
        try:
            build()
        except Exception:
            if not os.path.exists(ORACLE_LIB):
                raise
    lib = C.CDLL(ORACLE_LIB)
    vp = C.c_void_p
    lib.armada_oracle_round_schedule.argtypes = [C.POINTER(abi.RoundInput), C.POINTER(abi.RoundOutput), C.POINTER(abi.RoundStats)]
    lib.armada_oracle_round_schedule.restype = C.c_int32
    lib.armada_oracle_last_error.restype = C.c_char_p
    lib.armada_oracle_drf_cost.argtypes = [C.c_uint32, abi.i64p, abi.f64p, abi.i64p]
    lib.armada_oracle_drf_cost.restype = C.c_double
    lib.armada_oracle_nodedb_create.argtypes = [C.POINTER(abi.RoundInput), C.POINTER(vp)]
    lib.armada_oracle_nodedb_create.restype = C.c_int32
    lib.armada_oracle_nodedb_destroy.argtypes = [vp]
    lib.armada_oracle_nodedb_destroy.restype = None
    lib.armada_oracle_nodedb_schedule_many.argtypes = [vp, abi.u32p, C.c_uint32, abi.u8p, abi.u32p, abi.i32p, abi.i32p, abi.u8p]
    lib.armada_oracle_nodedb_schedule_many.restype = C.c_int32
    lib.armada_oracle_nodedb_dry_run.argtypes = [vp, abi.u32p, C.c_uint32, abi.u8p, abi.u32p]
    lib.armada_oracle_nodedb_dry_run.restype = C.c_int32
    lib.armada_oracle_nodedb_evict.argtypes = [vp, C.c_uint32]
    lib.armada_oracle_nodedb_evict.restype = C.c_int32
    lib.armada_oracle_nodedb_unbind.argtypes = [vp, C.c_uint32]
    lib.armada_oracle_nodedb_unbind.restype = C.c_int32
    lib.armada_oracle_nodedb_add_evicted.argtypes = [vp, C.c_uint32, C.c_int32]
    lib.armada_oracle_nodedb_add_evicted.restype = C.c_int32
    lib.armada_oracle_nodedb_get_alloc.argtypes = [vp, C.c_uint32, abi.i64p]
    lib.armada_oracle_nodedb_get_alloc.restype = C.c_int32
    lib.armada_oracle_nodedb_iterate.argtypes = [vp, C.c_uint32, C.c_int32, abi.i64p, abi.u32p, C.c_uint32, abi.u32p]
    lib.armada_oracle_nodedb_iterate.restype = C.c_int32
    lib.armada_oracle_node_index_key.argtypes = [C.c_uint32, C.c_uint64, abi.i64p, abi.i64p, C.c_uint64, C.c_int, abi.u8p]
    lib.armada_oracle_node_index_key.restype = C.c_int32
    _lib = lib
    return lib


def check(status: int):
    if status != abi.OK:
        raise abi.ArmadaError(status, load().armada_oracle_last_error().decode())


def round_schedule(inp: abi.RoundInput) -> RoundResult:
    lib = load()
    res = RoundResult(inp)
    check(lib.armada_oracle_round_schedule(C.byref(inp), C.byref(res.out), C.byref(res.stats)))
    return res


class OracleNodeDb:
    """NodeDb-level surface of the oracle (nodedb_test.go-style vectors)."""

    def __init__(self, inp: abi.RoundInput):
        self.lib = load()
        self.inp = inp
        self.h = C.c_void_p()
        check(self.lib.armada_oracle_nodedb_create(C.byref(inp), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.armada_oracle_nodedb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def schedule_many(self, jobs):
        n = len(jobs)
        ja = np.asarray(jobs, dtype=np.uint32)
        ok = C.c_uint8(0)
        node = np.full(n, abi.NONE, np.uint32)
        sa = np.zeros(n, np.int32)
        pa = np.zeros(n, np.int32)
        me = np.zeros(n, np.uint8)
        check(self.lib.armada_oracle_nodedb_schedule_many(
            self.h, ja.ctypes.data_as(abi.u32p), n, C.byref(ok), node.ctypes.data_as(abi.u32p),
            sa.ctypes.data_as(abi.i32p), pa.ctypes.data_as(abi.i32p), me.ctypes.data_as(abi.u8p)))
        return bool(ok.value), node, sa, pa, me

    def dry_run(self, jobs):
        """ScheduleManyWithTxn + Abort (what SubmitChecker does): (ok, nodes)"""
        n = len(jobs)
        ja = np.asarray(jobs, dtype=np.uint32)
        ok = C.c_uint8(0)
        node = np.full(n, abi.NONE, np.uint32)
        check(self.lib.armada_oracle_nodedb_dry_run(self.h, ja.ctypes.data_as(abi.u32p), n, C.byref(ok), node.ctypes.data_as(abi.u32p)))
        return bool(ok.value), node

    def evict(self, job):
        check(self.lib.armada_oracle_nodedb_evict(self.h, job))

    def unbind(self, job):
        check(self.lib.armada_oracle_nodedb_unbind(self.h, job))

    def add_evicted(self, job, index=-1):
        check(self.lib.armada_oracle_nodedb_add_evicted(self.h, job, index))

    def get_alloc(self, node):
        out = np.zeros((self.inp.num_priorities, self.inp.num_resources), np.int64)
        check(self.lib.armada_oracle_nodedb_get_alloc(self.h, node, out.ctypes.data_as(abi.i64p)))
        return out

    def iterate(self, row, priority, indexed_request):
        req = np.asarray(indexed_request, dtype=np.int64)
        out = np.zeros(max(1, self.inp.num_nodes), np.uint32)
        cnt = C.c_uint32(0)
        check(self.lib.armada_oracle_nodedb_iterate(self.h, row, priority, req.ctypes.data_as(abi.i64p),
                                                    out.ctypes.data_as(abi.u32p), len(out), C.byref(cnt)))
        return out[: cnt.value].tolist()
