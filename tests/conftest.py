import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


from armada_b200.model import RoundResult  # noqa: E402

RoundResult.FIRST_PASS_DEFAULT = True  # every parity test also compares the first-pass view (the inputs of queue_stats)
