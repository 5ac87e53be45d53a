"""CPU: libb200bo.so builds, loads, exports every symbol include/b200bo.h declares, and the Python ids
match the header.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200bo.h")


@pytest.fixture(scope="module")
def lib():
    from bayesian_optimization_b200 import build, _lib

    build.build()
    return _lib.load_library()


def header_text():
    return open(HEADER).read()


def test_every_declared_symbol_is_exported(lib):
    from bayesian_optimization_b200 import _lib

    txt = re.sub(r"/\*.*?\*/", "", header_text(), flags=re.S)
    declared = set(re.findall(r"\b(b200bo_[a-zA-Z0-9_]+)\s*\(", txt))
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/b200bo.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes prototypes and header disagree"


def test_constants_match_header():
    from bayesian_optimization_b200 import _lib

    defs = dict(re.findall(r"#define\s+B200BO_([A-Z0-9_]+)\s+\(?(-?\d+)\)?", header_text()))
    for k, v in defs.items():
        if hasattr(_lib, k):
            assert getattr(_lib, k) == int(v), k
    for k in ("CORR_RBF", "CORR_MATERN52", "CORR_GENEXP", "CORR_MATERN_NU", "MODE_NOISY", "TREND_LINEAR", "TREND_QUADRATIC",
              "ACQ_MGFI", "STATE_R", "FIT_REJECTED", "E_NODEVICE"):
        assert k in defs and hasattr(_lib, k)
    from oracle import gp_oracle as go

    for k in ("CORR_RBF", "CORR_MATERN12", "CORR_MATERN32", "CORR_MATERN52", "CORR_ABSEXP", "CORR_CUBIC", "CORR_GENEXP",
              "CORR_MATERN_NU", "TREND_CONSTANT", "TREND_LINEAR", "TREND_QUADRATIC", "MODE_NOISELESS", "MODE_NOISY", "MODE_NOISE_ESTIM", "ACQ_EI", "ACQ_PI", "ACQ_UCB", "ACQ_MGFI"):
        assert getattr(go, k) == int(defs[k]), k


def test_version_and_error_string(lib):
    assert lib.b200bo_version() >= 100
    assert isinstance(lib.b200bo_last_error(), bytes)


def test_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bayesian_optimization_b200 import B200BOError, Engine, _lib

    with pytest.raises(B200BOError) as e:
        Engine(0)
    assert e.value.code == _lib.E_NODEVICE
    h = ctypes.c_void_p()
    assert lib.b200bo_create(0, ctypes.byref(h)) == _lib.E_NODEVICE
    assert b"CUDA" in lib.b200bo_last_error() or b"device" in lib.b200bo_last_error()


def test_sass_is_sm100a():
    """the cubin inside the library targets sm_100a and the fp64 contraction is on the tensor pipe (DMMA)"""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from bayesian_optimization_b200 import _lib

    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
