"""CPU-side checks: the C-ABI library loads and exports every symbol that
include/ehb200.h declares; compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

import embeddinghub_b200 as ehb
from embeddinghub_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ehb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ehb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_native.SYMBOLS), set(names) ^ set(_native.SYMBOLS)
    assert ehb.lib().ehb_abi_version() == 2


def test_params_default_match_reference_defaults():
    p = _native.Params()
    ehb.lib().ehb_params_default(C.byref(p), 3)
    # index.cc:14-15 hnswlib defaults, index.h:21 init_cap
    assert (p.dim, p.M, p.ef_construction, p.ef_search, p.seed, p.capacity, p.metric) == (3, 16, 200, 10, 100, 128, 0)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ehb.EhbError) as e:
        ehb.NativeIndex(8)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "embeddinghub_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f in ("walk.cuh", "bruteforce.cu"), f


def test_rwlock_stress_host_only():
    """ehb::RwLock (writers preferred) under 8 spinning readers: exclusion holds and no writer starves.  Host code
    only, so it runs in the CPU tier."""
    import subprocess

    exe = os.path.join(ROOT, "tests", "cpp", "rwlock_stress")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/rwlock_stress not built (make)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
