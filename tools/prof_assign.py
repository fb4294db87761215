"""Counters of the node-assignment loop / the general loop (dev tooling; needs a GPU):
    python tools/prof_assign.py [scale] [config] [define]
`define` = ARMADA_BT_PROF (default: fresh nodes, candidate lookups, window refills, table GC inside
table_assign) or ARMADA_GL_PROF (general-loop iterations by outcome, level scans, batch overhead).
Builds armada_b200/libarmada_b200_prof.so with that define when it is missing or older than the sources."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from armada_b200 import abi, synth  # noqa: E402
from armada_b200.scheduler import DeviceRound  # noqa: E402

import subprocess  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
name = sys.argv[2] if len(sys.argv) > 2 else "C3"
define = sys.argv[3] if len(sys.argv) > 3 else "ARMADA_BT_PROF"
CSRC = os.path.join(ROOT, "armada_b200", "csrc")
PROF_LIB = os.path.join(ROOT, "armada_b200", "libarmada_b200_prof.so")
STAMP = PROF_LIB + ".define"
srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
stale = (not os.path.exists(PROF_LIB) or any(os.path.getmtime(f) > os.path.getmtime(PROF_LIB) for f in srcs)
         or not os.path.exists(STAMP) or open(STAMP).read().strip() != define)
if stale:
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                    "-Xcompiler", "-pthread", "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-diag-suppress", "550",
                    "-D" + define, "-o", PROF_LIB, os.path.join(CSRC, "armada_round.cu")], check=True)
    open(STAMP, "w").write(define)
lib = C.CDLL(PROF_LIB)
abi.declare_prototypes(lib)
r = synth.scaled(name, scale) if scale < 1.0 else {"C3": synth.config_c3, "C2": synth.config_c2, "C4": synth.config_c4}[name]()
with DeviceRound(0, lib=lib) as dev:
    dev.upload(r.to_input())
    for _ in range(3):
        st = dev.run()
    d = list(st.batch_debug)
    c = list(st.batch_cycles)
    print("placements", st.placements, "batch_cycles", c)
    print("debug", d, "iters", st.loop_iterations, "probes", st.probes, "rescans", st.tree_rescans, "pass_ms", st.schedule_pass_ms, "phase", list(st.phase_cycles))
    if define == "ARMADA_GL_PROF":
        print("general iterations ok", d[0], "cycles each", d[1] / max(d[0], 1), "| failed", d[2], "cycles each", d[3] / max(d[2], 1))
        print("level scans", d[5], "cycles each", d[4] / max(d[5], 1), "| main loop cycles", d[6], "| in run_batch", d[7])
        sys.exit(0)
    print("fresh placements", d[0], "cycles/fresh", d[1] / max(d[0], 1))
    print("candidate lookups", d[3], "cycles/lookup", d[2] / max(d[3], 1), "cursor chunks", d[6])
    print("cycles per cursor chunk", d[7] / max(d[6], 1))
    print("table gc", d[5], "assign loop cycles", d[4])
