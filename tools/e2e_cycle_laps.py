"""Where an end-to-end cycle of 8 pools spends its time (dev tooling; needs a GPU): per pool thread the wall-clock
span of upload / run / download inside one PoolCycle.schedule_cycle-like fan-out."""
import os
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from armada_b200 import synth  # noqa: E402
from armada_b200.model import RoundResult  # noqa: E402
from armada_b200.scheduler import DeviceRound  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
inputs = [synth.config_c3(seed=synth.SEED + p).to_input() for p in range(P)]
devs = [DeviceRound(0) for _ in range(P)]
results = [RoundResult(inputs[p]) for p in range(P)]
spans = [None] * P


def go(i, t0):
    a = time.perf_counter()
    devs[i].upload(inputs[i])
    b = time.perf_counter()
    st = devs[i].run()
    c = time.perf_counter()
    devs[i].download(results[i])
    d = time.perf_counter()
    spans[i] = tuple(round((x - t0) * 1e3, 1) for x in (a, b, c, d)) + (round(st.device_ms, 1),)


for rep in range(3):
    t0 = time.perf_counter()
    ts = [threading.Thread(target=go, args=(i, t0)) for i in range(P)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    total = (time.perf_counter() - t0) * 1e3
    print(f"cycle {rep}: {total:.1f} ms; per pool (start, upload done, run done, download done, device_ms):")
    for s in spans:
        print("   ", s)
