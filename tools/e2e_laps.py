"""Upload / run / download wall-clock of ONE pool through the C ABI with host buffers, with the upload's phase laps
(ARMADA_TIME_UPLOAD=1).  Dev tooling; needs a GPU."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from armada_b200 import synth
from armada_b200.scheduler import DeviceRound
r = synth.config_c3()
inp = r.to_input()
with DeviceRound(0) as dev:
    dev.schedule(inp)
    os.environ["ARMADA_TIME_UPLOAD"] = "1"
    t = time.perf_counter(); dev.upload(inp); t1 = time.perf_counter(); st = dev.run(); t2 = time.perf_counter(); res = dev.download(); t3 = time.perf_counter()
    print("upload %.1f ms run %.1f ms download %.1f ms" % ((t1 - t) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
