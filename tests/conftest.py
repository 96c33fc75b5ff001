import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    _ensure_library()


def _ensure_library():
    """A fresh checkout has no lib/libqsb200.so (built artefacts are git-ignored): build it once with nvcc so that
    the ABI tests (CPU) and the parity tests (GPU) exercise the real library.  No nvcc -> the tests that need the
    library fail loudly, which is the intended behaviour (there is no fallback)."""
    import importlib.util
    import shutil
    lib = os.path.join(ROOT, "quantum-simulation_b200", "lib", "libqsb200.so")
    if os.path.exists(lib) or not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        return
    spec = importlib.util.spec_from_file_location("_qsb_build", os.path.join(ROOT, "quantum-simulation_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build(force=True)


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


_cache = {}


def load_golden(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return _cache[name]


def golden_mpo(key, dtype=None, device="cpu"):
    """MPO site tensors stored by tests/golden/make_golden.py (distinct tensors + index list)."""
    z = load_golden("mpo")
    out = []
    for u in z[f"{key}/index"]:
        t = torch.from_numpy(z[f"{key}/t{int(u)}"])
        if dtype is not None:
            t = t.to(dtype)
        out.append(t.to(device).contiguous())
    return out


def ncon_layout(t):
    """The strided layout the reference's own env updates emit: shape (W, chi, chi), strides (chi, 1, W*chi)."""
    return t.permute(2, 0, 1).contiguous().permute(1, 2, 0)


def rel_err(a, b):
    a = a.detach().cpu().reshape(-1)
    b = b.detach().cpu().reshape(-1)
    return float(torch.linalg.norm(a - b) / max(float(torch.linalg.norm(b)), 1e-300))
