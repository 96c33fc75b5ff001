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

    with contextlib.redirect_stdout(io.StringIO()):
        m1 = qsb.MPS(args.L, 2, args.chi, "cpu", dtype, seed=1).to(dev)
        m2 = qsb.MPS(args.L, 2, args.chi, "cpu", dtype, seed=1).to(dev)
        d1 = qsb.DMRG(m1, mpo, device=dev, dtype=dtype)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        Es, _ = d1(sched(), dispon=0)
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    qsb.clear_workspaces()
    torch.cuda.empty_cache()
    d2 = parallel.ShardedDMRG(m2, mpo, device=dev, dtype=dtype)
    d2.profile = args.profile
    dist.barrier()
    t0 = time.perf_counter()
    Ed, _ = d2(sched(), dispon=0)
    torch.cuda.synchronize()
    dist.barrier()
    t_shard = time.perf_counter() - t0
    Es, Ed = np.array(Es), np.array(Ed)
    rel = np.abs(Ed - Es) / np.abs(Es)
    if rank == 0:
        print(json.dumps({"case": f"Heisenberg L={args.L} chi={args.chi} {args.dtype}, 1 sweep ({len(Es)} bond updates)",
                          "world": world, "sweep_s_single_gpu": t_single, "sweep_s_sharded": t_shard,
                          "speedup": t_single / t_shard, "E_end_single": float(Es[-1]), "E_end_sharded": float(Ed[-1]),
                          "rel_dev_end_of_sweep": float(rel[-1]), "max_rel_dev_updates_after_20": float(rel[20:].max()),
                          "bond_dims_equal": d1.history["bond_dims"] == d2.history["bond_dims"], "comm": d2.comm,
                          "phase_seconds_rank0": d2.timers if args.profile else None}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
