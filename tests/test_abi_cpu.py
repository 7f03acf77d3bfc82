"""CPU: the C-ABI library loads, exports every symbol include/e4s_b200.h declares, mirrors struct E4SConv
exactly, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from e4s2024_b200.build import build
    build()
    from e4s2024_b200 import _lib
    return _lib


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "e4s_b200.h")).read()
    declared = set(re.findall(r"\b(e4s_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = lib.lib()
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)


def test_struct_mirror(lib):
    assert lib.lib().e4s_sizeof_conv() == ctypes.sizeof(lib.E4SConv)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    sm = ctypes.c_int(0)
    rc = lib.lib().e4s_device_info(ctypes.byref(sm), None, None)
    assert rc != 0 and b"CUDA" in lib.lib().e4s_last_error()
    with pytest.raises(lib.E4SError):
        lib.nchw_to_nhwc(torch.zeros(1, 3, 4, 4))          # CPU tensor -> loud failure
    from e4s2024_b200.stylegan2.op import upfirdn2d
    with pytest.raises(RuntimeError):
        upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))


def test_arg_validation_messages(lib):
    p = lib.E4SConv()
    rc = lib.lib().e4s_conv_f32(ctypes.byref(p), None)
    assert rc == -1 and b"null" in lib.lib().e4s_last_error()
