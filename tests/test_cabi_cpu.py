"""CPU-only checks of the drop-in boundary: the library builds/loads and exports every symbol
include/cetpick.h declares; argument validation that needs no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from cet_pick_b200 import build, _lib
    build.build()
    return _lib


def test_header_symbols_all_exported(L):
    hdr = open(os.path.join(ROOT, "include", "cetpick.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cetpick_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = L.lib()
    for name in declared:
        assert getattr(lib, name) is not None


def test_test_hooks_live_in_the_test_library_only(L):
    """include/cetpick_test.h declares the per-kernel hooks and probes; libcetpick_test_sm100a.so exports them (and the
    whole product ABI), the product library exports none of them."""
    hdr = open(os.path.join(ROOT, "include", "cetpick_test.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cetpick_[a-z0-9_]+)\s*\(", hdr))
    assert declared and declared == set(L.TEST_SIGNATURES), declared ^ set(L.TEST_SIGNATURES)
    assert not (declared & set(L.SIGNATURES))
    tl, pl = L.test_lib(), L.lib()
    for name in declared:
        assert getattr(tl, name) is not None
        assert not hasattr(pl, name), f"{name} leaked into the product library"
    for name in L.SIGNATURES:
        assert getattr(tl, name) is not None


def test_synthetic_data_helpers_are_outside_the_product_package():
    assert not os.path.exists(os.path.join(ROOT, "cet_pick_b200", "synth.py"))
    for dp, _, fs in os.walk(os.path.join(ROOT, "cet_pick_b200")):
        for f in fs:
            if f.endswith(".py"):
                assert "synthdata" not in open(os.path.join(dp, f)).read(), os.path.join(dp, f)


def test_version_and_strerror(L):
    lib = L.lib()
    assert lib.cetpick_version() == 1
    assert lib.cetpick_strerror(0) == b"ok"
    assert b"workspace" in lib.cetpick_strerror(L.ERR_WORKSPACE)


def test_argument_validation_without_gpu(L):
    lib = L.lib()
    n = C.c_size_t(0)
    assert lib.cetpick_decode_workspace_bytes(8, 16, 16, 10, C.byref(n)) == 0 and n.value > 0
    assert lib.cetpick_decode_workspace_bytes(8, 16, 16, 8 * 16 * 16 + 1, C.byref(n)) == L.ERR_BAD_ARG
    assert lib.cetpick_decode_workspace_bytes(0, 16, 16, 1, C.byref(n)) == L.ERR_BAD_ARG
    assert lib.cetpick_decode_f32(None, 1, 1, 1, 1, 3, 1, 1, None, None, None, None, 0, None) == L.ERR_BAD_ARG
    h = C.c_void_p()
    assert lib.cetpick_unet_create(C.byref(h), 4, 32, 32) == 0
    w = C.c_size_t(0)
    assert lib.cetpick_unet_workspace_bytes(h, 128, 512, 512, 0, C.byref(w)) == 0
    assert w.value > 128 * 256 * 256 * 32 * 2 * 3
    # forward before finalize is a state error, never a silent fallback
    assert lib.cetpick_unet_forward(h, 1, 4, 32, 32, 1, 0, None, None, 0, None) == L.ERR_STATE
    assert lib.cetpick_unet_create(C.byref(C.c_void_p()), 4, 64, 32) == L.ERR_UNSUPPORTED
    lib.cetpick_unet_destroy(h)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under cet_pick_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "cet_pick_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_no_cpu_fallback_on_cpu_tensors(L):
    import torch
    from cet_pick_b200.models.decode import tomo_decode
    with pytest.raises(RuntimeError):
        tomo_decode(torch.zeros(1, 1, 4, 8, 8), K=4)
