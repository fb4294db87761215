"""Repeat rounds on the device to shake out timing-dependent protocol bugs: python tools/stress_gpu.py [scale] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from armada_b200 import synth  # noqa: E402
from armada_b200.scheduler import DeviceRound  # noqa: E402

name = sys.argv[3] if len(sys.argv) > 3 else "C3"
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
r = synth.scaled(name, scale) if scale < 1.0 else {"C3": synth.config_c3, "C2": synth.config_c2, "C4": synth.config_c4}[name]()
inp = r.to_input()
ref = None
with DeviceRound(0) as dev:
    dev.upload(inp)
    t0 = time.time()
    for i in range(reps):
        st = dev.run()
        got = dev.download()
        if ref is None:
            ref = got
        else:
            bad = got.diff(ref)
            if bad:
                print(f"rep {i}: NON-DETERMINISTIC result: {bad[:3]}")
                sys.exit(2)
    print(f"{name}@{scale}: {reps} identical rounds, placements={st.placements}, {1e3 * (time.time() - t0) / reps:.1f} ms/round incl. download")
