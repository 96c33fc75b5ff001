#!/usr/bin/env python3
"""Stage the UNMODIFIED reference package under baseline/_ref/ so that it travels to the GPU box.

    python baseline/stage_reference.py            (run in the build container, where /root/reference exists)

The reference (ToelUl/quantum-simulation) is a pure-Python tree without setup.py / pyproject.toml, so it cannot be
pip-installed; this recipe copies the package directory byte for byte (``quantum_simulation/**/*.py``) to
``baseline/_ref/quantum_simulation`` and writes a sha256 manifest next to it.  ``baseline/_ref/`` is git-ignored (the
history stays free of reference sources) but not gpurun-ignored, exactly like the built ``.so`` files.  Consumers:

  * ``bench.py --impl reference``   times the reference's own ``ApplyMPO('ncon')`` on the host cores
                                    (``cpu_baseline.kind = "reference"``; the oracle port is the fallback)
  * ``tests/test_gpu_dropin_reference.py``   drives the reference's own ``DMRG`` / ``lanczos_ground_state`` on
                                    ``device='cuda'`` through the binding of ``integration/qsb200_binding.py``

``import_reference()`` is the loader both use: it installs the four import stubs the package needs
(matplotlib / IPython are not in the image) and imports it from baseline/_ref only.
"""
import hashlib
import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = "/root/reference"
PKG = "quantum_simulation"


def stage(src: str = SRC, dest: str = DEST) -> int:
    """Copy the package; returns the number of files staged (0 if the source tree is not present)."""
    root = os.path.join(src, PKG)
    if not os.path.isdir(root):
        return 0
    out = os.path.join(dest, PKG)
    if os.path.isdir(out):
        shutil.rmtree(out)
    manifest = []
    for dirpath, dirnames, filenames in os.walk(root):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        for fn in sorted(filenames):
            if not fn.endswith(".py"):
                continue
            a = os.path.join(dirpath, fn)
            rel = os.path.relpath(a, src)
            b = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(b), exist_ok=True)
            shutil.copyfile(a, b)
            manifest.append((hashlib.sha256(open(b, "rb").read()).hexdigest(), rel))
    for extra in ("LICENSE", "LICENSE.md", "LICENSE.txt"):
        if os.path.exists(os.path.join(src, extra)):
            shutil.copyfile(os.path.join(src, extra), os.path.join(dest, extra))
    with open(os.path.join(dest, "MANIFEST.sha256"), "w") as f:
        f.write(f"# staged unmodified from {src} by baseline/stage_reference.py\n")
        for h, rel in sorted(manifest, key=lambda t: t[1]):
            f.write(f"{h}  {rel}\n")
    return len(manifest)


def staged() -> bool:
    return os.path.exists(os.path.join(DEST, PKG, "dmrg.py"))


def import_reference():
    """``import quantum_simulation`` from baseline/_ref (never from /root/reference: that path does not exist on the
    GPU box).  Raises ImportError if the tree has not been staged."""
    if not staged():
        raise ImportError("reference not staged: run `python baseline/stage_reference.py` in the build container")
    for name in ("matplotlib", "matplotlib.pyplot", "IPython", "IPython.display"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["IPython"].display = sys.modules["IPython.display"]
    for attr in ("display", "Math"):
        if not hasattr(sys.modules["IPython.display"], attr):
            setattr(sys.modules["IPython.display"], attr, lambda *a, **k: None)
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    sys.dont_write_bytecode = True
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module(PKG)
    if not os.path.abspath(mod.__file__).startswith(os.path.abspath(DEST)):
        raise ImportError(f"quantum_simulation resolved to {mod.__file__}, not to the staged copy")
    return mod


if __name__ == "__main__":
    n = stage()
    print(f"staged {n} files under {DEST}" if n else f"{SRC} not present: nothing staged")
