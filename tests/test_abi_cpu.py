"""CPU: the C-ABI library loads, exports every symbol include/meshgen_b200.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "meshgen_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from reinforcementlearning4meshgeneration_b200 import _lib
    _lib.build()
    L = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"libmeshgen_b200.so does not export {n}"
    assert set(_lib.SYMBOLS) == set(names)
    assert b"sm_100a" in L.mg_version()


def test_sass_is_sm100a_and_uses_bulk_copy():
    import shutil
    import subprocess
    from reinforcementlearning4meshgeneration_b200 import _lib
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out, "the vertex ring must be staged with a bulk async copy (cp.async.bulk -> UBLKCP)"
    assert "DFMA" in out or "DADD" in out


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU container")
    from reinforcementlearning4meshgeneration_b200 import BatchedBoudaryEnv, _lib
    L = _lib.load()
    h = C.c_void_p()
    assert L.mg_create(C.byref(h), 0, 4, 64) != 0
    assert b"no CPU fallback" in L.mg_last_error(None)
    with pytest.raises(RuntimeError):
        BatchedBoudaryEnv([[(0, 0), (0, 1), (1, 1), (1, 0)]], num_envs=1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "reinforcementlearning4meshgeneration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, f"{f} mentions the oracle"
