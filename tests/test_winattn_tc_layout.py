"""CPU replay of the shared-memory layouts / UMMA descriptors of the tcgen05 window-attention kernels.

tests/native/winattn_tc_layout_check.cpp includes the kernels' own host/device layout header
(fiber_b200/csrc/window_tc_layout.cuh), writes a shared-memory image through the writer functions and
fetches every tcgen05.mma operand back through an independent model of the canonical UMMA layouts.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_winattn_tc_layout_replay(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "tcl_check")
    src = os.path.join(ROOT, "tests", "native", "winattn_tc_layout_check.cpp")
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "layouts OK" in r.stdout


def test_winattn_tc_option_roundtrip():
    """The default routing is the TMA-fed tcgen05 generation (15); the option round-trips through the C-ABI and -1 restores it."""
    from fiber_b200 import lib
    if os.environ.get("FIBER_WINATTN_TC"):
        pytest.skip("FIBER_WINATTN_TC set in the environment")
    assert lib.get_option("winattn_tc") == 15
    lib.set_option("winattn_tc", 0)
    try:
        assert lib.get_option("winattn_tc") == 0
    finally:
        lib.set_option("winattn_tc", -1)
    assert lib.get_option("winattn_tc") == 15
    assert lib.get_option("winattn_tc_launches") == 0
    with pytest.raises(RuntimeError):
        lib.get_option("no_such_option")
