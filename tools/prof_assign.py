"""Counters of the node-assignment loop (build with -DARMADA_BT_PROF into libarmada_b200_prof.so):
python tools/prof_assign.py [scale] [config]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from armada_b200 import abi, synth  # noqa: E402
from armada_b200.scheduler import DeviceRound  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
name = sys.argv[2] if len(sys.argv) > 2 else "C3"
lib = C.CDLL(os.path.join(ROOT, "armada_b200", "libarmada_b200_prof.so"))
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
    if d[0]: print("per touched placement:", [round(x / d[0], 1) for x in d[1:5]])
    print("fresh placements", d[0], "cycles/fresh", d[1] / max(d[0], 1))
    print("candidate lookups", d[3], "cycles/lookup", d[2] / max(d[3], 1), "cursor chunks", d[6])
    print("cycles per cursor chunk", d[7] / max(d[6], 1))
    print("table gc", d[5], "assign loop cycles", d[4])
