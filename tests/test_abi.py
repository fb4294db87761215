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



def test_header_is_plain_c99_and_a_c_caller_links(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 (what cgo's C compiler sees), a C translation unit that
    fills the structs and calls an entry point links against the product library, and the field offsets a C
    compiler computes are the ones the ctypes mirror uses."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    fields = [("ArmadaRoundInput", "queued_order", abi.RoundInput.queued_order.offset),
              ("ArmadaRoundInput", "gang_uniformity_label", abi.RoundInput.gang_uniformity_label.offset),
              ("ArmadaRoundInput", "floating_limit", abi.RoundInput.floating_limit.offset),
              ("ArmadaRoundOutput", "job_reason_first_pass", abi.RoundOutput.job_reason_first_pass.offset),
              ("ArmadaRoundOutput", "num_result_preempted", abi.RoundOutput.num_result_preempted.offset),
              ("ArmadaRoundStats", "batch_debug", abi.RoundStats.batch_debug.offset)]
    checks = "\n".join(f'  if (offsetof({t}, {f}) != {off}) {{ printf("{t}.{f} %zu != {off}\\n", offsetof({t}, {f})); bad = 1; }}' for t, f, off in fields)
    src.write_text(f'''#include <stddef.h>
#include <stdio.h>
#include "armada_b200.h"
int main(void) {{
  int bad = 0;
{checks}
  ArmadaRoundInput in = {{0}};
  in.abi_version = ARMADA_ABI_VERSION;
  if (armada_abi_version() != in.abi_version) bad = 1;
  if (armada_abi_sizeof(0) != sizeof(ArmadaRoundInput)) bad = 1;
  return bad;
}}
''')
    exe = tmp_path / "caller"
    if not os.path.exists(abi.PRODUCT_LIB_PATH):
        pytest.skip("libarmada_b200.so not built")
    libdir = os.path.dirname(abi.PRODUCT_LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-larmada_b200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
