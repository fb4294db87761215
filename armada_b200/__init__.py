"""armada_b200 — B200-native implementation of Armada's scheduling-round hot path.

Only what the path needs lives here:
  csrc/        hand-written CUDA (sm_100a) + the C ABI (include/armada_b200.h)
  abi.py       ctypes mirror of the C ABI
  model.py     host-side mirror of the reference's scheduler-internal types → SoA flattening
  scheduler.py PreemptingQueueScheduler-shaped entry point backed by the CUDA library
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
