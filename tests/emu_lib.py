"""TEST TOOLING ONLY: builds armada_b200/csrc/armada_round.cu with g++ against
tools/simt_emu/cuda_emu.h (a deterministic single-threaded SIMT emulator) so the kernel *logic*
can be run through the parity suite without a GPU.  Not a product path, not a fallback: nothing
under armada_b200/ loads this library."""
from __future__ import annotations

import ctypes as C
import fcntl
import os
import subprocess

from armada_b200 import abi
from armada_b200.scheduler import DeviceRound

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tools", "simt_emu")
EMU_LIB = os.path.join(EMU_DIR, "libarmada_b200_emu.so")
CSRC = os.path.join(ROOT, "armada_b200", "csrc")
_lib = None


def build(force: bool = False) -> None:
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".inc", ".h"))]
    srcs += [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(ROOT, "include", "armada_b200.h")]
    def stale() -> bool:
        return not os.path.exists(EMU_LIB) or any(os.path.getmtime(p) > os.path.getmtime(EMU_LIB) for p in srcs)

    if not (force or stale()):
        return
    # pytest-xdist workers race for the build: one builds (to a temporary name, renamed into place), the others wait
    with open(EMU_LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or stale():
            tmp = f"{EMU_LIB}.{os.getpid()}.tmp"
            subprocess.run(
                ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DARMADA_EMU", "-x", "c++",
                 os.path.join(CSRC, "armada_round.cu"), "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", EMU_DIR,
                 "-o", tmp, "-lpthread"], check=True)
            os.replace(tmp, EMU_LIB)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(EMU_LIB)
        abi.declare_prototypes(_lib)
    return _lib


def emu_round() -> DeviceRound:
    return DeviceRound(0, lib=load())
