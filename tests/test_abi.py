"""The C ABI of include/armada_b200.h: every declared entry point is exported by the product
library (built by nvcc for sm_100a; no compute call is made here — there is no GPU), and the
ctypes mirror in armada_b200/abi.py has the struct sizes the compiled libraries have."""
import ctypes as C
import os
import re

import pytest

import emu_lib
import oracle_lib
from armada_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "armada_b200.h")


def declared_entry_points():
    src = open(HEADER).read()
    names = re.findall(r"^\s*(?:int32_t|uint32_t|const char\*)\s+(armada_\w+)\s*\(", src, flags=re.M)
    assert len(names) >= 9
    return sorted(set(names))


def test_python_symbol_list_matches_the_header():
    assert sorted(abi.PRODUCT_SYMBOLS) == declared_entry_points()


def test_product_library_exports_every_entry_point():
    if not os.path.exists(abi.PRODUCT_LIB_PATH):
        pytest.skip("libarmada_b200.so not built (run __graft_entry__.build())")
    lib = C.CDLL(abi.PRODUCT_LIB_PATH)
    for sym in declared_entry_points():
        assert hasattr(lib, sym), f"{sym} missing from {abi.PRODUCT_LIB_PATH}"
    lib.armada_abi_version.restype = C.c_uint32
    assert lib.armada_abi_version() == abi.ABI_VERSION
    lib.armada_abi_sizeof.restype = C.c_uint32
    lib.armada_abi_sizeof.argtypes = [C.c_uint32]
    assert lib.armada_abi_sizeof(0) == C.sizeof(abi.RoundInput)
    assert lib.armada_abi_sizeof(1) == C.sizeof(abi.RoundOutput)
    assert lib.armada_abi_sizeof(2) == C.sizeof(abi.RoundStats)


def test_struct_sizes_match_the_emulated_build_and_the_oracle():
    lib = emu_lib.load()
    assert lib.armada_abi_sizeof(0) == C.sizeof(abi.RoundInput)
    assert lib.armada_abi_sizeof(1) == C.sizeof(abi.RoundOutput)
    assert lib.armada_abi_sizeof(2) == C.sizeof(abi.RoundStats)
    ora = oracle_lib.load()
    ora.armada_oracle_abi_sizeof.restype = C.c_uint32
    ora.armada_oracle_abi_sizeof.argtypes = [C.c_uint32]
    for which, t in enumerate((abi.RoundInput, abi.RoundOutput, abi.RoundStats)):
        assert ora.armada_oracle_abi_sizeof(which) == C.sizeof(t)


def test_product_fails_loudly_without_a_device():
    """No CPU fallback: on a box without a GPU the product refuses to create a round."""
    if not os.path.exists(abi.PRODUCT_LIB_PATH):
        pytest.skip("libarmada_b200.so not built")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from armada_b200.scheduler import DeviceRound
    with pytest.raises(abi.ArmadaError) as ei:
        DeviceRound(0)
    assert ei.value.status in (abi.E_NO_DEVICE, abi.E_CUDA)

