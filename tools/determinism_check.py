#!/usr/bin/env python3
"""Which operation makes a seeded GPU sweep non-repeatable?  (three runs of the seeded C3
sweep gave three end-of-sweep energies, 2e-6 apart.)

Part 1 runs every building block twice on identical inputs and compares the results BITWISE.
Part 2 runs the same short DMRG twice while hashing every intermediate of every bond step (theta, Lanczos vector,
U / S / Vh of the split, the new environment) and reports the first quantity that differs.

    python tools/determinism_check.py [--chi 512] [--L 16]          (one GPU; prints JSON lines)
"""
import argparse
import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantum_simulation_b200 as qsb           # noqa: E402
from quantum_simulation_b200 import svd as qsvd  # noqa: E402
from quantum_simulation_b200 import dmrg as qdmrg  # noqa: E402

DEV = "cuda"


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def mpo(key, dtype):
    z = np.load(os.path.join(ROOT, "tests", "golden", "mpo.npz"))
    return [torch.from_numpy(z[f"{key}/t{int(u)}"]).to(dtype).contiguous().to(DEV) for u in z[f"{key}/index"]]


def h(t):
    """Bit-level fingerprint: SHA-1 of the bytes for small tensors; for large ones two wrap-around int64 sums over
    the raw 64-bit words computed on the device (a full-size sweep hashes ~200 GB of intermediates)."""
    t = t.detach().contiguous()
    if t.numel() * t.element_size() > (1 << 22) and t.is_cuda and t.element_size() in (8, 16):
        w = torch.view_as_real(t).reshape(-1).view(torch.int64) if t.is_complex() else t.reshape(-1).view(torch.int64)
        idx = torch.arange(w.numel(), device=w.device, dtype=torch.int64)
        return f"{int(w.sum().item()) & 0xffffffffffff:012x}{int((w * (2 * idx + 1)).sum().item()) & 0xffffff:06x}"
    return hashlib.sha1(t.cpu().numpy().tobytes()).hexdigest()[:12]


def twice(name, fn, reps=3):
    outs = []
    for _ in range(reps):
        r = fn()
        r = r if isinstance(r, (tuple, list)) else (r,)
        outs.append([h(x) for x in r])
    same = all(o == outs[0] for o in outs)
    print(json.dumps({"op": name, "bitwise_repeatable": same, "hashes": outs if not same else outs[0]}), flush=True)
    return same


def part1(chi):
    dt = torch.float64
    g = torch.Generator(device=DEV).manual_seed(5)
    n = 2 * chi
    a = torch.randn(n, n, dtype=dt, device=DEV, generator=g)
    # spectrum like a DMRG theta: fast decay, many near-null directions
    U0, _ = torch.linalg.qr(a)
    s = torch.exp(-torch.arange(n, dtype=dt, device=DEV) * (30.0 / n))
    theta = (U0 * s) @ torch.linalg.qr(a.T)[0]
    G = theta @ theta.T
    twice("torch.matmul (cuBLAS) Gram", lambda: theta @ theta.T)
    twice("torch.linalg.eigh (cuSOLVER syevd)", lambda: torch.linalg.eigh(G))
    twice("torch.linalg.qr tall (geqrf+orgqr)", lambda: torch.linalg.qr(theta[:, :chi]))
    twice("torch.linalg.svd small 64x64 (default driver)", lambda: torch.linalg.svd(theta[:64, :64].contiguous()))
    twice("torch.linalg.svd 512x512 (default driver)", lambda: torch.linalg.svd(theta[:512, :512].contiguous()))
    twice("torch.linalg.cholesky", lambda: torch.linalg.cholesky(G + torch.eye(n, dtype=dt, device=DEV)))
    twice("svd.two_site_split (gram path)", lambda: qsvd.two_site_split(theta, chi, 0.0)[:3])
    twice("qsb gemm core", lambda: qsb._lib.gemm(theta, theta.T))
    Ms = mpo("heis_f64_L20", dt)
    L = torch.randn(5, chi, chi, dtype=dt, device=DEV, generator=g)
    R = torch.randn(5, chi, chi, dtype=dt, device=DEV, generator=g)
    L = L + L.transpose(1, 2)
    R = R + R.transpose(1, 2)
    psi = torch.randn(chi * 4 * chi, dtype=dt, device=DEV, generator=g)
    op = qsb.ApplyMPO()
    twice("qsb heff_apply", lambda: op(psi, L, Ms[9], Ms[10], R).clone())
    A = torch.randn(chi, 2, chi, dtype=dt, device=DEV, generator=g)
    twice("qsb env_left", lambda: qsb.LeftEnvUpdate()(L, Ms[9], A))
    twice("qsb env_right", lambda: qsb.RightEnvUpdate()(Ms[9], R, A))
    twice("qsb lanczos_ground_state",
          lambda: qsb.lanczos_ground_state(psi, op, (L, Ms[9], Ms[10], R), num_restarts=2, krylov_dim=4)[0].clone())
    twice("MPS built on the GPU (torch.rand + QR sweep), seed 1",
          lambda: [t.detach() for t in quiet(qsb.MPS, 12, 2, chi, DEV, dt, 1).B])


def part_big(n=4096):
    """The library calls of a chi = n/2 bond step at full size: repeatable inside a process?  (Run the tool twice and
    diff the logs for the cross-process answer: the hashes are printed.)"""
    dt = torch.float64
    g = torch.Generator(device=DEV).manual_seed(11)
    a = torch.randn(n, n, dtype=dt, device=DEV, generator=g)
    U0, _ = torch.linalg.qr(a)
    s = torch.exp(-torch.arange(n, dtype=dt, device=DEV) * (30.0 / n))
    theta = (U0 * s) @ torch.linalg.qr(a.T)[0]
    G = qsb._lib.gemm(theta, theta.T)
    G = 0.5 * (G + G.T)
    k = n // 2
    twice(f"BIG qsb gemm Gram {n}", lambda: qsb._lib.gemm(theta, theta.T), 2)
    twice(f"BIG torch.matmul {n}", lambda: theta @ theta.T, 2)
    twice(f"BIG torch.linalg.eigh {n}", lambda: torch.linalg.eigh(G), 2)
    twice(f"BIG torch.linalg.qr {n}x{k}", lambda: torch.linalg.qr(theta[:, :k]), 2)
    Gk = G[:k, :k] + torch.eye(k, dtype=dt, device=DEV)
    twice(f"BIG cholesky_ex {k}", lambda: torch.linalg.cholesky_ex(Gk)[0], 2)
    Lc = torch.linalg.cholesky(Gk)
    twice(f"BIG solve_triangular {k}x{n}", lambda: torch.linalg.solve_triangular(Lc, theta[:k], upper=False), 2)
    twice(f"BIG two_site_split {n}", lambda: qsvd.two_site_split(theta, k, 0.0)[:3], 2)
    twice("BIG einsum lsc,ck->lsk", lambda: torch.einsum("lsc,ck->lsk", theta.view(k, 2, n), a[:, :k].contiguous()), 2)
    twice(f"BIG MPS built on the GPU, L=14 chi={k // 2}, seed 1",
          lambda: [t.detach() for t in quiet(qsb.MPS, 14, 2, k // 2, DEV, dt, 1).B], 2)
    twice("BIG torch.rand with a device generator",
          lambda: torch.rand(1000, 1000, generator=torch.Generator(device=DEV).manual_seed(1), device=DEV, dtype=dt), 2)


class Tap:
    """Records a hash of every intermediate of every bond step of one DMRG run."""

    def __init__(self):
        self.log = []

    def install(self):
        tap = self
        self._split = qdmrg.two_site_split
        self._lan = qdmrg.lanczos_ground_state
        self._mm = qdmrg._matmul2d
        self._upL = qdmrg.EnvironmentManager.update_L
        self._upR = qdmrg.EnvironmentManager.update_R

        def split(theta, *a, **k):
            tap.log.append(("split_in", h(theta)))
            U, S, Vh, keep = tap._split(theta, *a, **k)
            tap.log += [("U", h(U)), ("S", h(S)), ("Vh", h(Vh))]
            return U, S, Vh, keep

        def lan(x0, *a, **k):
            tap.log.append(("theta0", h(x0)))
            v, e = tap._lan(x0, *a, **k)
            tap.log.append(("lanczos_vec", h(v)))
            return v, e

        def upL(self_, p, A):
            tap._upL(self_, p, A)
            tap.log.append(("envL", h(self_.L_cache[p + 1])))

        def upR(self_, p, B):
            tap._upR(self_, p, B)
            tap.log.append(("envR", h(self_.R_cache[p])))

        qdmrg.two_site_split, qdmrg.lanczos_ground_state = split, lan
        qdmrg.EnvironmentManager.update_L, qdmrg.EnvironmentManager.update_R = upL, upR

    def remove(self):
        qdmrg.two_site_split, qdmrg.lanczos_ground_state = self._split, self._lan
        qdmrg.EnvironmentManager.update_L, qdmrg.EnvironmentManager.update_R = self._upL, self._upR


def part2(chi, L, mps_device, seed_global=None):
    dt = torch.float64
    Ms = mpo("heis_f64_L100", dt)
    Ms = Ms[:L // 2] + Ms[-(L - L // 2):]
    runs = []
    for _ in range(2):
        tap = Tap()
        tap.install()
        try:
            mps = quiet(qsb.MPS, L, 2, chi, mps_device, dt, 1).to(DEV)
            sw = qsb.Sweeps(1)
            sw.set_schedule(maxdim=[chi], cutoff=[0.0], krylov_dim=[4], noise=[0.0])
            dm = quiet(qsb.DMRG, mps, Ms, device=DEV, dtype=dt)
            if seed_global is not None:
                torch.manual_seed(seed_global)
            E, _ = quiet(dm, sw, dispon=0, maxit=2)
        finally:
            tap.remove()
        runs.append((tap.log, np.array(E)))
    (la, Ea), (lb, Eb) = runs
    first = next((i for i, (x, y) in enumerate(zip(la, lb)) if x != y), None)
    rel = np.abs(Ea - Eb) / np.abs(Ea)
    digest = hashlib.sha1(repr(la).encode()).hexdigest()[:16]       # compare across PROCESSES by diffing two logs
    print(json.dumps({"dmrg_twice": f"Heisenberg L={L} chi={chi} f64, 1 sweep, MPS built on {mps_device}, global RNG "
                                    + ("unseeded" if seed_global is None else f"seeded ({seed_global})"),
                      "records": len(la), "log_digest_run1": digest,
                      "first_records": [list(r) for r in la[:6]], "first_differing_record": None if first is None else [first, la[first][0]],
                      "context": None if first is None else [r[0] for r in la[max(0, first - 3): first + 2]],
                      "energies_bitwise_equal": bool(np.array_equal(Ea, Eb)), "max_rel_energy_dev": float(rel.max()),
                      "end_of_sweep": [float(Ea[-1]), float(Eb[-1])]}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=512)
    ap.add_argument("--L", type=int, default=16)
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--seed-global", type=int, default=None,
                    help="torch.manual_seed before each sweep (a zero start vector is re-seeded from the global RNG)")
    ap.add_argument("--sweep-only", action="store_true", help="only the twice-in-one-process DMRG with hashed intermediates")
    args = ap.parse_args()
    if args.big:
        part_big()
        part2(args.chi, args.L, DEV)
        sys.exit(0)
    if args.sweep_only:
        part2(args.chi, args.L, DEV, args.seed_global)
        sys.exit(0)
    part1(args.chi)
    part2(min(args.chi, 256), args.L, "cpu")
    part2(min(args.chi, 256), args.L, DEV)
