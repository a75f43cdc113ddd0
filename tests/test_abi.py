"""The C-ABI shared library: builds, loads and exports every symbol include/rpx.h declares
(no compute calls: this runs without a GPU), and the numpy/ctypes layout mirrors match."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from raypier_optics_b200 import _abi as A
from raypier_optics_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "rpx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rpx_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "librpx.so has not been built (python -c 'import __graft_entry__ as g; g.build()')"
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_functions() if not hasattr(L, s)]
    assert not missing, missing
    L.rpx_abi_version.restype = ctypes.c_int
    assert L.rpx_abi_version() == A.RPX_ABI_VERSION


def test_sm100a_cubin_embedded():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the library must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.load()
    ctx = ctypes.c_void_p()
    rc = L.rpx_init(0, ctypes.byref(ctx))
    assert rc == -5  # RPX_ERR_NODEVICE
    assert b"no CPU fallback" in L.rpx_last_error(None)
    from raypier_optics_b200.engine import Engine
    with pytest.raises(_lib.RpxError):
        Engine(0)


def test_record_layouts():
    assert A.ray_dtype.itemsize == 188 and A.gausslet_dtype.itemsize == 668
    assert A.ray_dtype.fields['length'][1] == 144 and A.ray_dtype.fields['wavelength_idx'][1] == 168
    assert A.gausslet_dtype.fields['para_rays'][1] == 188
    assert A.face_dtype.fields['tolerance'][1] == 40 and A.face_dtype.fields['p'][1] == 48


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under raypier_optics_b200/ may reference it."""
    pkg = os.path.join(ROOT, "raypier_optics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "rpx_oracle" not in text, f
