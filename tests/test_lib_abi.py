"""CPU-side checks of the C-ABI library: it loads, and exports every symbol include/uncltmo_b200.h declares."""
import ctypes
import os
import re
import subprocess

from uncltmo_b200 import _lib


def test_header_declares_entry_points():
    names = _lib.declared_symbols()
    for must in ("uncl_conv3x3_tc", "uncl_conv3x3_simt", "uncl_gcn_knn_aggregate", "uncl_frame_normalise_pad",
                 "uncl_tiles_blend", "uncl_percentile_pair", "uncl_last_error", "uncl_arch"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m uncltmo_b200.build` first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.declared_symbols():
        assert hasattr(lib, name), name


def test_arch_is_sm100a_only():
    lib = _lib.lib()
    assert lib.uncl_arch() == b"sm_100a"
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode == 0:
        archs = set(re.findall(r"sm_\d+a?", out.stdout))
        assert archs == {"sm_100a"}, archs
    ptx = subprocess.run(["cuobjdump", "-lptx", _lib.LIB_PATH], capture_output=True, text=True)
    assert "sm_" not in ptx.stdout  # SASS only: no PTX that could JIT for another arch


def test_no_cpu_path():
    import pytest
    import torch
    with pytest.raises(RuntimeError):
        _lib.call("uncl_maxpool2", torch.zeros(8), 0, None, 0, 0, torch.zeros(8), 0, 1, 8, 2, 2, 0)
