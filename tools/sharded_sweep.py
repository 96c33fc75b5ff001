#!/usr/bin/env python3
"""Sweep-level multi-GPU measurement: one full two-site DMRG sweep of a Heisenberg chain through ShardedDMRG on all
ranks vs the single-GPU DMRG (every rank runs it redundantly on its own GPU), same CPU-seeded MPS.
  torchrun --nproc-per-node N tools/sharded_sweep.py [--L 40] [--chi 1024] [--dtype f64]"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import quantum_simulation_b200 as qsb           # noqa: E402
from quantum_simulation_b200 import parallel    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=40)
    ap.add_argument("--chi", type=int, default=1024)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    rank, world, local = bench.dist_setup(0)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.complex128
    full = bench.load_mpo("heis_f64_L100" if dtype == torch.float64 else "heis_c128_L100", dtype)
    mpo = [full[0], full[1]] + [full[50]] * (args.L - 4) + [full[-2], full[-1]]      # uniform chain of L sites
    mpo = [m.to(dev) for m in mpo]

    def sched():
        sw = qsb.Sweeps(1)
        sw.set_schedule(maxdim=[args.chi], cutoff=[0.0], krylov_dim=[4], noise=[0.0])
        return sw

    def make_mps():
        """Random right-canonical MPS built on the GPU and made identical on every rank (rank 0's tensors)."""
        with contextlib.redirect_stdout(io.StringIO()):
            m = qsb.MPS(args.L, 2, args.chi, dev, dtype, seed=1)
        for plist in (m.A, m.B, m.sWeight):
            for t in plist:
                buf = torch.view_as_real(t.data) if t.is_complex() else t.data
                dist.broadcast(buf, src=0)
        return m

    m1, m2 = make_mps(), make_mps()
    with contextlib.redirect_stdout(io.StringIO()):
        d1 = qsb.DMRG(m1, mpo, device=dev, dtype=dtype)
        d1(sched(), dispon=0)                       # sweep 1: warm-up (allocations, cuSOLVER handles)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        Es, _ = d1(sched(), dispon=0)               # sweep 2: timed
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    qsb.clear_workspaces()
    torch.cuda.empty_cache()
    d2 = parallel.ShardedDMRG(m2, mpo, device=dev, dtype=dtype)
    d2(sched(), dispon=0)                           # sweep 1: warm-up (NCCL communicators, symmetric arena)
    d2.profile = args.profile
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    Ed, _ = d2(sched(), dispon=0)                   # sweep 2: timed
    torch.cuda.synchronize()
    dist.barrier()
    t_shard = time.perf_counter() - t0
    Es, Ed = np.array(Es), np.array(Ed)
    rel = np.abs(Ed - Es) / np.abs(Es)
    if rank == 0:
        print(json.dumps({"case": f"Heisenberg L={args.L} chi={args.chi} {args.dtype}, SECOND sweep timed ({len(Es)} bond updates)",
                          "world": world, "sweep_s_single_gpu": t_single, "sweep_s_sharded": t_shard,
                          "speedup": t_single / t_shard, "E_end_single": float(Es[-1]), "E_end_sharded": float(Ed[-1]),
                          "rel_dev_end_of_sweep": float(rel[-1]), "max_rel_dev_all_updates": float(rel.max()),
                          "bond_dims_equal": d1.history["bond_dims"] == d2.history["bond_dims"], "comm": d2.comm,
                          "phase_seconds_rank0": d2.timers if args.profile else None}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
