"""The C-ABI library must load without a GPU and export exactly what include/rsb200.h declares."""

import ctypes
import os
import re

import pytest

from robosat_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "rsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "librsb200.so does not export %s" % n
    # and the ctypes table binds the same set
    assert sorted(_lib.SIGNATURES) == names


def test_struct_layout_matches_header():
    # sizes follow from the C declarations (8-byte aligned, no packing pragmas)
    assert ctypes.sizeof(_lib.ConvSrc) == 56
    assert ctypes.sizeof(_lib.ConvSeg) == 16
    assert _lib.ConvDesc.srcs.offset == 8 and _lib.ConvDesc.segs.offset == 8 + 4 * 56 + 4
    # the library reports the sizes it was compiled with; _lib.load() refuses to bind on a mismatch
    sizes = (ctypes.c_int32 * 4)()
    _lib.load().rsb_abi_layout(sizes)
    assert list(sizes) == [ctypes.sizeof(_lib.ConvSrc), ctypes.sizeof(_lib.ConvSeg), ctypes.sizeof(_lib.ConvDesc), ctypes.sizeof(_lib.RowConvDesc)]
    assert _lib.RSB_MAX_SEGS == 16 and _lib.RSB_MAX_SRCS == 4  # keep in sync with include/rsb200.h
    assert ctypes.sizeof(_lib.ConvDesc) % 8 == 0


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.rsb_version() >= 200
    assert isinstance(_lib.last_error(), str)


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.RsbError):
        _lib.require_device()
    from robosat_b200 import synth
    from robosat_b200.engine import UNetEngine

    with pytest.raises(_lib.RsbError):
        UNetEngine(synth.make_state_dict(2), 2, 1, 64, 64, device="cpu")


def test_descriptor_validation_runs_before_the_device_check():
    """Argument errors are reported as RSB_E_INVALID on any host (validation never touches the device or dereferences the
    pointers); a VALID descriptor then fails with RSB_E_NODEVICE here -- there is no CPU execution path behind a plan."""
    import ctypes
    import sys

    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conv_cases

    lib = _lib.load()
    plan = ctypes.c_void_p()
    case = conv_cases.conv_case("1x1", 1, 16, 16, 64, 64, "cpu", relu=False, bias=False)
    assert lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)) == -3          # valid, but no sm_100 device
    buf = torch.zeros(1 << 16, dtype=torch.float32)
    case.desc.stats, case.desc.stats_bytes = buf.data_ptr(), 64
    assert lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)) == -1          # stats buffer too small
    assert "stats" in _lib.last_error()
    case.desc.stats_bytes = buf.numel() * 4
    assert lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)) == -3
    res = conv_cases.conv_case("1x1", 1, 16, 16, 64, 64, "cpu", relu=False, bias=False, residual=True)
    res.desc.stats, res.desc.stats_bytes = buf.data_ptr(), buf.numel() * 4
    assert lib.rsb_conv_plan_create(ctypes.byref(res.desc), ctypes.byref(plan)) == -1           # stats + residual
    case.desc.stats = None
    case.desc.block_n = 48
    assert lib.rsb_conv_plan_create(ctypes.byref(case.desc), ctypes.byref(plan)) == -1           # not a tile width
    # host codec entry points work without a device; the stream inflate rejects garbage
    out = (ctypes.c_uint8 * 16)()
    assert lib.rsb_zlib_inflate(b"\x78\x9c" + b"\xff" * 20, 22, out, 16) == -1
    assert lib.rsb_pack_weights1(None, None, None, 8, None) == -1
