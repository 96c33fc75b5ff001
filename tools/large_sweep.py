#!/usr/bin/env python3
"""north_star at SWEEP level: one full two-site DMRG sweep of BASELINE config 4 (J1-J2 Heisenberg cylinder 8 x 6 = 48
sites, J2/J1 = 0.5, FSM MPO bonds up to 32, float64) at chi = 4096 through ShardedDMRG on all GPUs of the node.

One GPU cannot run this configuration: the 2 (N + 1) environments alone are 32 * 4096^2 * 8 B = 4.3 GB each, 0.42 TB
together; row-sharded over 8 GPUs they are 52 GB per GPU.  No CPU reference exists at this size
either, so the result is checked for self-consistency: <psi|H|psi> of the final MPS by an independent environment
contraction on one GPU (observables.energy_expectation) against the last Lanczos energy.

    torchrun --nproc-per-node 8 tools/large_sweep.py [--chi 4096] [--model j1j2]
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import quantum_simulation_b200 as qsb           # noqa: E402
from quantum_simulation_b200 import parallel, svd as qsvd    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=4096)
    ap.add_argument("--model", default="j1j2", choices=["j1j2", "hubbard", "heis"])
    ap.add_argument("--sites", type=int, default=0, help="use only the first/last sites of the MPO (0 = the full chain)")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--shard-mps", action="store_true", help="keep the MPS itself row-sharded (ShardedDMRG(shard_mps=True))")
    args = ap.parse_args()
    rank, world, local = bench.dist_setup(0)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    kf, kc, site, desc = bench.MODELS[args.model]
    dtype = torch.complex128 if args.model == "hubbard" else torch.float64
    mpo = [m.to(dev) for m in bench.load_mpo(kf, dtype)]
    if args.sites:
        h = args.sites // 2
        mpo = mpo[:h] + mpo[-(args.sites - h):]
    L, d = len(mpo), mpo[0].shape[2]
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        mps = qsb.MPS(L, d, args.chi, dev, dtype, seed=1, rescale_carry=True)
    for plist in (mps.A, mps.B, mps.sWeight):                       # identical tensors on every rank (rank 0's)
        for t in plist:
            dist.broadcast(torch.view_as_real(t.data) if t.is_complex() else t.data, src=0)
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    sw = qsb.Sweeps(1)
    sw.set_schedule(maxdim=[args.chi], cutoff=[0.0], krylov_dim=[4], noise=[0.0])
    dm = parallel.ShardedDMRG(mps, mpo, device=dev, dtype=dtype, shard_mps=args.shard_mps)
    dm.profile = args.profile
    before = dict(qsvd.stats)
    torch.manual_seed(1)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    E, _ = dm(sw, dispon=0, maxit=2)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mem = torch.tensor([torch.cuda.max_memory_allocated() / 1e9], dtype=torch.float64, device=dev)
    dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    bd = dm.history["bond_dims"]
    res = {"case": f"{desc}, {L} sites, d={d}, chi={args.chi}, {'c128' if dtype.is_complex else 'f64'}: one full two-site "
                   f"sweep through ShardedDMRG on {world} GPUs (MPO bonds up to {max(m.shape[1] for m in mpo)})",
           "world": world, "mps_row_sharded": bool(args.shard_mps),
           "mps_elements_stored_per_rank": int(sum(x.numel() for x in mps.A) + sum(x.numel() for x in mps.B)),
           "sweep_s": float(t.item()), "bond_updates": len(E), "updates_at_full_chi": int(sum(b == args.chi for b in bd)),
           "mps_init_on_gpu_s": t_init, "E_end_of_sweep": float(E[-1]), "E_mid": float(E[len(E) // 2 - 1]),
           "peak_memory_gb_max_over_ranks": float(mem.item()), "comm": dm.comm,
           "split_paths": {k: qsvd.stats[k] - before.get(k, 0) for k in ("sketch_accepted", "sketch_rejected", "sketch_skipped", "gram", "svd")},
           "discarded_weight_max": float(max(dm.history["discarded_weights"])),
           "phase_seconds_rank0": dm.timers if args.profile else None}
    if not args.no_check:
        # independent check on rank 0: <H> of the final (replicated) MPS by plain left-environment contraction
        if args.shard_mps:
            dm.gather_mps()
        del dm
        qsb.clear_workspaces()
        torch.cuda.empty_cache()
        if rank == 0:
            t0 = time.perf_counter()
            res["E_by_environments_one_gpu"] = float(qsb.energy_expectation(mps, mpo))
            res["energy_check_abs_dev"] = abs(res["E_by_environments_one_gpu"] - res["E_end_of_sweep"])
            res["energy_check_s"] = time.perf_counter() - t0
        dist.barrier()
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
