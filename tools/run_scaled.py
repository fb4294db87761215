"""Run one scaled C3 round on the device (for ncu captures): python tools/run_scaled.py [scale] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from armada_b200 import synth  # noqa: E402
from armada_b200.scheduler import DeviceRound  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
name = sys.argv[3] if len(sys.argv) > 3 else "C3"
r = synth.scaled(name, scale) if scale < 1.0 else {"C3": synth.config_c3, "C2": synth.config_c2}[name]()
inp = r.to_input()
with DeviceRound(0) as dev:
    dev.upload(inp)
    for _ in range(reps):
        st = dev.run()
    print(f"{name}@{scale}: nodes={inp.num_nodes} jobs={inp.num_jobs} placements={st.placements} probes={st.probes} "
          f"device_ms={st.device_ms:.3f} pass_ms={st.schedule_pass_ms:.3f} launches={st.gpu_launches}")
