"""The C-ABI library loads and exports every symbol include/medgp_cuda.h declares; without a
GPU the only call made is medgp_cuda_create, which must refuse (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "medgp_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(medgp_cuda_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from medgp_b200 import api
    assert header_symbols() == sorted(api.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from medgp_b200 import api
    lib = api.load_library()
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    from medgp_b200 import api
    with pytest.raises(api.MedgpError):
        api.Context(2, 2, 2)
    lib = api.load_library()
    h = ctypes.c_void_p()
    assert lib.medgp_cuda_create(ctypes.byref(h), 0, 0) == -4   # MEDGP_ERR_NODEVICE


def test_host_front_ends_link_the_cuda_library_only():
    """the shipped binaries must depend on libmedgp_cuda.so and not on the oracle"""
    import subprocess
    for name in ("main_one_train", "main_one_test", "main_cohort_train"):
        path = os.path.join(ROOT, "medgp_b200", "host", name)
        if not os.path.exists(path):
            pytest.skip("host binaries not built")
        out = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
        assert "libmedgp_cuda.so" in out and "oracle" not in out
