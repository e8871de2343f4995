"""The C-ABI library loads and exports every symbol include/cvs_b200.h declares; nothing computes without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cvs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvs_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(cvs):
    from cudavideostream_b200.api import SIGNATURES
    lib = ctypes.CDLL(cvs.library_path())
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cvs_b200.h but not exported by libcvs_b200.so"
        assert n in SIGNATURES, f"{n} has no ctypes signature in cudavideostream_b200/api.py"
    assert set(SIGNATURES) <= set(names)


def test_cudacore_shim_symbols_are_exported(cvs):
    # the four members of diff::cuda::CUDACore (server/include/kernels.cuh:38-41), Itanium-mangled
    lib = ctypes.CDLL(cvs.library_path())
    for sym in ["_ZN4diff4cuda8CUDACoreC1EPhRNS_5utils5matszEPfiS2_S5_",
                "_ZN4diff4cuda8CUDACore12alloc_arraysEPPhS3_S3_PPiii",
                "_ZN4diff4cuda8CUDACore9exec_coreEPhS2_RNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEEPjPi",
                "_ZN4diff4cuda8CUDACore11chunkt_sizeEv"]:
        assert hasattr(lib, sym), sym


def test_abi_version_and_defaults(cvs):
    from cudavideostream_b200.api import _Config
    lib = cvs.load_library()
    assert lib.cvs_abi_version() == 1
    cfg = _Config()
    lib.cvs_config_default(ctypes.byref(cfg))
    # defaults = the reference's compile-time switches (server/include/common.h:4-18, threads.cpp:37-38)
    assert (cfg.width, cfg.height, cfg.threshold, cfg.mode, cfg.noise_filter, cfg.ksize) == (1920, 1080, 20, 0, 0, 3)


def test_no_cpu_fallback(cvs):
    # on a box without an sm_100 device every compute entry point must refuse loudly
    if cvs.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(cvs.CVSError) as e:
        cvs.Stream(16, 16, np.zeros(3 * 16 * 16, dtype=np.uint8))
    assert e.value.status == 6  # CVS_ERR_NODEVICE
    with pytest.raises(cvs.CVSError):
        cvs.alloc_host(1024)
    with pytest.raises(cvs.CVSError):
        cvs.filters.heat_map(16, 16, 16, 4, 4)


def test_invalid_arguments_are_rejected_before_touching_the_device(cvs):
    from cudavideostream_b200.api import _Config
    lib = cvs.load_library()
    h = ctypes.c_void_p()
    assert lib.cvs_create(None, ctypes.byref(h)) == 1  # CVS_ERR_INVALID
    cfg = _Config()
    lib.cvs_config_default(ctypes.byref(cfg))
    cfg.width = 0
    assert lib.cvs_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
    assert b"frame size" in lib.cvs_last_error()
    cfg.width, cfg.mode = 64, 9
    assert lib.cvs_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
    assert lib.cvs_destroy(None) == 0


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under cudavideostream_b200/ may import, include, link or call it
    pkg = os.path.join(ROOT, "cudavideostream_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                text = open(path).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "liboracle" not in text, f"{f} loads the oracle library"
            elif f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(path, errors="replace").read()
                code = re.sub(r"//.*?$|/\*.*?\*/", "", text, flags=re.S | re.M)
                assert not re.search(r'#include\s+[<"][^>"]*oracle', code), f"{f} includes the oracle"
                assert not re.search(r"\borc_\w+\s*\(", code), f"{f} calls the oracle"
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libcvs_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
