"""bench.py prints ONE JSON line with the contract's keys: the reference arm on the CPU (here, small chi) and our
arm on the GPU (marked gpu)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout=600):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_contract_cpu():
    d = _run(["--impl", "reference", "--chi", "64", "--steps", "2", "--warmup", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "GFLOP/s" and d["higher_is_better"] is True and d["value"] > 0
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "quantum_simulation", "dmrg.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_our_arm_contract_gpu():
    d = _run(["--chi", "256", "--steps", "3", "--warmup", "3"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    for k in ("clocks", "roofline", "gpu_launches", "lanczos", "env_update"):
        assert k in d
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1.2 and r["peak"] > 10
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["gpu_launches"] == 3 * 4                 # prepare + GEMM + mix + GEMM per step
    assert d["executed"]["identity_channels"] == {"L": 4, "R": 0}          # canonical-gauge environments (FSM layout)
    assert d["executed"]["tflops"] < d["tflops"] and r["channels_executed"]["step1"] == 4
    assert "exchange" not in d["config"]              # both arms print the same config
    assert d["e2e"]["h2d_bytes_per_step"] == 256 * 4 * 256 * 8 and d["e2e"]["value"] < d["value"] * 1.05
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] > 0
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
