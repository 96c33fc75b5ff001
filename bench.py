#!/usr/bin/env python3
"""Benchmark of the two-site DMRG hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--chi 2048] [--dtype f64|c128]
                    [--model heis|j1j2|hubbard|tfim] [--exchange auto|fused|nccl] [--no-extras]

Workload (config.workload): BASELINE config 3 -- Heisenberg chain L=100, chi=2048, the bulk bond (sites 50|51),
MPO bond W=5, d=2.  A *step* is one H_eff matvec  out = L . M1 . M2 . R . theta  (reference
quantum_simulation/dmrg.py:178-228, the operator Lanczos applies 8x per bond update).

  value        whole-job matvec throughput in GFLOP/s with every input resident in HBM, counted with the dense
               ncon-order flop count the reference executes (qsb_heff_flops) -- an EFFECTIVE rate: the
               environments are in the canonical gauge a DMRG sweep produces (one identity channel in L and in R for a
               finite-state-machine MPO), and the kernels skip the GEMM slices of those channels; ``executed`` carries
               the flops actually executed, and the roofline / utilisation numbers are computed from THOSE
               (``--dense-env`` benchmarks fully random environments instead: nothing to skip)
  e2e          the same metric through the reference-facing call ``ApplyMPO.forward`` with theta AND the result in
               pinned HOST memory: upload of theta + matvec + download of the result inside the timed region (the
               call pipelines them: column-block upload gating the step-1 GEMM, row-block download overlapped);
               N > 1: every rank uploads its shard and downloads its rows
  roofline     the dominant kernel (the FP64 tensor-core GEMM of step 1), timed with CUDA events recorded on the
               launching stream inside the timed region (qsb_heff_set_stage_events); peak = FP64 tensor-pipe
               issue rate measured in this run by the register-resident DMMA probe (MEASURED_PEAKS.json has no
               FP64 entry); cuBLAS DGEMM 8192^3 and the same matvec run by torch/cuBLAS on this GPU are reported
               beside it
  cpu_baseline the oracle (torch-CPU restatement of the reference's ncon path) timed on the host cores
  lanczos / env_update   the other hot-path kernels: achieved HBM GB/s of the fused vector ops, env-update GFLOP/s

N > 1: the matvec is sharded over the left bond index a' (rows of L and of the result); every step exchanges the
theta shards -- by default fused into the matvec (P2P copy-engine pushes + arrival-gated step-1 GEMM, no NCCL on the
data path; ``--exchange nccl`` = all-gather then matvec) -- and runs the local row block: strong scaling of the same
chi.  ``--model`` selects the MPO bond of another BASELINE config (j1j2 = config 4, hubbard = config 5, tfim = 2).
``--impl reference`` times the oracle on the host cores (rank 0 only) for the same config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "H_eff matvec GFLOP/s (two-site DMRG bulk bond; default BASELINE config 3: Heisenberg L=100, chi=2048)"
UNIT = "GFLOP/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chi", type=int, default=2048)
    ap.add_argument("--dtype", default="f64", choices=["f64", "c128"])
    ap.add_argument("--model", default="heis", choices=sorted(MODELS),
                    help="which BASELINE config's MPO bond to run (default: config 3, the headline)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "nccl"],
                    help="N>1: how theta shards are exchanged (auto = fused P2P push if available, else NCCL)")
    ap.add_argument("--dense-env", action="store_true",
                    help="fully random L and R (no identity channels): the round-1 workload, nothing is skipped")
    ap.add_argument("--serial-e2e", action="store_true",
                    help="N > 1: time the e2e leg as copy -> fused matvec -> copy (the r01 path) instead of forward_host")
    ap.add_argument("--no-sweep", action="store_true", help="skip the measured full DMRG sweep of the extras leg")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / lanczos / env / library legs")
    return ap.parse_args()


def load_mpo(key, dtype):
    z = np.load(os.path.join(ROOT, "tests", "golden", "mpo.npz"))
    return [torch.from_numpy(z[f"{key}/t{int(u)}"]).to(dtype).contiguous() for u in z[f"{key}/index"]]


MODELS = {   # model -> (MPO key for f64, for c128, left site of the bond, description)
    "heis": ("heis_f64_L100", "heis_c128_L100", 50, "BASELINE config 3: Heisenberg chain L=100"),
    "j1j2": ("j1j2_f64_8x6", "j1j2_f64_8x6", 20, "BASELINE config 4: J1-J2 Heisenberg cylinder 8x6 (J2/J1=0.5)"),
    "hubbard": ("hubbard_f64_L64", "hubbard_f64_L64", 30, "BASELINE config 5: Fermi-Hubbard chain via Jordan-Wigner, d=4, L=64"),
    "tfim": ("tfim_c128_L100_g0.5", "tfim_c128_L100_g0.5", 50, "BASELINE config 2: transverse-field Ising chain L=100"),
}
_MODEL = "heis"
_DENSE_ENV = False


def fsm_identity_channels(mpo, site):
    """Which channel of L[site] and of R[site+2] is the identity for a canonical MPS?  Structural propagation through
    the MPO chain (the reference's FSM layout, mpo/auto_mpo/finite_state_machine.py:662-700): the boundary
    environments are ones(W,1,1) (dmrg.py:349-351); channel w' of L[p+1] is A^H A = 1 iff column w' of M[p] holds a
    single non-zero block, that block is the d x d identity, and it comes from an identity channel of L[p].  Mirror
    image from the right.  Returns (idL, idR), -1 where there is none.  (tests/test_gpu_contract.py checks the same
    channels NUMERICALLY on environments built from a canonical MPS.)"""
    def step(M, ids, left):
        Wi, Wo = (M.shape[0], M.shape[1]) if left else (M.shape[1], M.shape[0])
        blk = (lambda i, o: M[i, o]) if left else (lambda i, o: M[o, i])
        eye = torch.eye(M.shape[2], dtype=M.dtype)
        out = set()
        for o in range(Wo):
            src = [i for i in range(Wi) if bool((blk(i, o) != 0).any())]
            if len(src) == 1 and src[0] in ids and torch.equal(blk(src[0], o), eye):
                out.add(o)
        return out
    idsL = set(range(mpo[0].shape[0])) if mpo[0].shape[0] == 1 else set()
    for p in range(site):
        idsL = step(mpo[p].cpu(), idsL, True)
    idsR = set(range(mpo[-1].shape[1])) if mpo[-1].shape[1] == 1 else set()
    for p in range(len(mpo) - 1, site + 1, -1):
        idsR = step(mpo[p].cpu(), idsR, False)
    return (min(idsL) if idsL else -1), (min(idsR) if idsR else -1)


def make_inputs(chi, dtype, device, rows=None, row0=0):
    """Seeded synthetic bulk-bond inputs: real MPO site tensors of the model, random L/R/theta.
    Generated per row so that any row shard sees exactly the rows of the global problem.  Unless --dense-env, the
    environments are in the gauge a sweep produces: the FSM's "nothing yet" channel of L and "all done" channel of R
    are identity matrices (A^H A = 1 for a canonical MPS); every other channel is dense random."""
    kf, kc, site, _ = MODELS[_MODEL]
    mpo = load_mpo(kf if dtype == torch.float64 else kc, dtype)
    idL, idR = (-1, -1) if _DENSE_ENV else fsm_identity_channels(mpo, site)
    M1, M2 = mpo[site].to(device), mpo[site + 1].to(device)
    W = M1.shape[0]
    g = torch.Generator(device="cpu").manual_seed(0)
    rows = chi if rows is None else rows
    gl = torch.Generator(device=device).manual_seed(1000 + row0)
    L = torch.randn(W, rows, chi, dtype=dtype, device=device, generator=gl) / chi ** 0.5
    gr = torch.Generator(device=device).manual_seed(1)
    R = torch.randn(M2.shape[1], chi, chi, dtype=dtype, device=device, generator=gr) / chi ** 0.5
    gt = torch.Generator(device=device).manual_seed(2)
    theta = torch.randn(chi * M1.shape[3] * M2.shape[3] * chi, dtype=dtype, device=device, generator=gt)
    theta /= torch.linalg.norm(theta)
    if idL >= 0:
        L[idL].zero_()
        r = torch.arange(rows, device=device)
        L[idL, r, r + row0] = 1.0
    if idR >= 0:
        R[idR] = torch.eye(chi, dtype=dtype, device=device)
    return L, M1, M2, R, theta


def hermitian_envs(L, R):
    """Hermitian versions of square environments for the Lanczos legs, identity channels kept as they are."""
    def herm(env):
        eye = torch.eye(env.shape[1], dtype=env.dtype, device=env.device)
        out = env + env.conj().transpose(1, 2)
        for w in range(env.shape[0]):
            if torch.equal(env[w], eye):
                out[w] = eye
        return out
    return herm(L), herm(R)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


# ----------------------------------------------------------------------------------------------
# reference arm: the oracle (torch-CPU restatement of the reference's ncon path) on the host cores
# ----------------------------------------------------------------------------------------------
def reference_apply_mpo():
    """(callable, kind): the UNMODIFIED reference's ``ApplyMPO('ncon')`` when it has been staged under baseline/_ref
    (baseline/stage_reference.py; kind "reference"), else the oracle port of the same tensordot chain (kind "port")."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_stage_reference", os.path.join(ROOT, "baseline", "stage_reference.py"))
        st = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(st)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            st.import_reference()
            from quantum_simulation.dmrg import ApplyMPO
        mod = ApplyMPO("ncon")
        return (lambda psi, L, M1, M2, R, out=None: mod(psi, L, M1, M2, R, out=out)), "reference"
    except Exception:
        from oracle import dmrg_oracle as orc
        return orc.apply_mpo, "port"


def cpu_matvec_time(chi, dtype, reps, rows=None, threads=None):
    from oracle import dmrg_oracle as orc
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    L, M1, M2, R, theta = make_inputs(chi, dtype, "cpu", rows=rows)
    flops = orc.matvec_flops(L, M1, M2, R)
    out = torch.empty(L.shape[1] * M1.shape[2] * M2.shape[2] * R.shape[1], dtype=dtype)
    apply_mpo, kind = reference_apply_mpo()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        apply_mpo(theta, L, M1, M2, R, out=out)
        ts.append(time.perf_counter() - t0)
    return ts, flops, threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = torch.float64 if args.dtype == "f64" else torch.complex128
    cores = os.cpu_count() or 1
    # one full-size matvec decides the bounded sample: a block of `rows` left-bond rows of the same problem
    ts, flops_full, threads, kind = cpu_matvec_time(args.chi, dtype, 1)
    budget = 200.0
    total_steps = args.steps + args.warmup
    frac = min(1.0, budget / max(ts[0] * total_steps, 1e-9))
    rows = max(64, int(args.chi * frac) // 64 * 64) if frac < 1.0 else args.chi
    rows = min(rows, args.chi)
    ts, flops, _, kind = cpu_matvec_time(args.chi, dtype, total_steps, rows=rows)
    timed = ts[args.warmup:]
    sec = sum(timed) / len(timed)
    val = flops / sec / 1e9
    sample = (f"rows [0,{rows}) of the left bond (of {args.chi}) of the same chi={args.chi} matvec per step, "
              f"{len(timed)} timed steps, "
              + ("the unmodified reference's ApplyMPO('ncon') (baseline/_ref)" if kind == "reference" else
                 "oracle port of the reference's ncon-order tensordots")
              + f" on torch CPU ({torch.__config__.parallel_info().splitlines()[0]})")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, world):
    kf, kc, site, desc = MODELS[args.model]
    mpo = load_mpo(kf if args.dtype == "f64" else kc, torch.float64 if args.dtype == "f64" else torch.complex128)
    M1, M2 = mpo[site], mpo[site + 1]
    return {"workload": f"{desc}, chi={args.chi}, bulk bond {site}|{site + 1}, MPO bonds "
                        f"{M1.shape[0]}-{M1.shape[1]}-{M2.shape[1]}, d={M1.shape[2]}, "
                        f"{'float64' if args.dtype == 'f64' else 'complex128'}; one H_eff matvec per step",
            "chi": args.chi, "d": int(M1.shape[2]), "W": int(M1.shape[0]), "hamiltonian": args.model,
            "environments": ("dense random" if args.dense_env else
                             "canonical gauge: identity channel in L and R (as a sweep produces), other channels dense random"),
            "parallelism": "single GPU" if world == 1 else f"left-bond row shards x{world}",
            "l2_policy": "inputs + intermediates per step (>1 GB at chi=2048) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def measured_traffic(args, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu
    --set full capture of this round (profiles/r02_ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep);
    None when no capture matches this configuration."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    key = f"{args.model}/chi{args.chi}/{args.dtype}/n{world}/{'dense' if args.dense_env else 'canonical'}"
    try:
        z = json.load(open(path))
        e = z[key]
        return float(e["dram_bytes"]), f"{path.replace(ROOT + os.sep, '')}[{key}]: {e.get('source', '')}"
    except Exception:
        return None, f"no ncu capture committed for {key}"


def time_region(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def run_ours(args):
    import torch.distributed as dist
    import quantum_simulation_b200 as qsb
    from quantum_simulation_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.complex128
    es = 8 if dtype == torch.float64 else 16
    chi = args.chi
    assert chi % world == 0
    rows = chi // world
    L, M1, M2, R, theta = make_inputs(chi, dtype, dev, rows=rows, row0=rank * rows)
    # what EnvironmentManager does after every environment update: test the channels for the identity (one small
    # kernel + a W-double read-back) and remember the answer on the tensor; ApplyMPO.forward picks it up
    _lib.tag_identity(L, rank * rows)
    _lib.tag_identity(R)
    ident = (_lib.identity_of(L)[0], _lib.identity_of(R)[0], rank * rows)
    n = theta.numel()
    n_loc = rows * M1.shape[2] * M2.shape[2] * chi
    op = qsb.ApplyMPO("ncon")
    out = torch.empty(n_loc, dtype=dtype, device=dev)
    flops_local = _lib.heff_flops(theta, L, M1, M2, R)                       # dense ncon-order count ("effective")
    flops_exec_local = _lib.heff_flops_executed(theta, L, M1, M2, R, ident)  # what the kernels execute
    flops_total = flops_local * world

    theta_shard = theta.view(world, -1)[rank].clone() if world > 1 else None
    exchange = "none"
    fused = None
    if world > 1:
        exchange = "nccl all-gather of theta, then the matvec"
        if args.exchange != "nccl":
            try:           # all-gather fused into the matvec: P2P pushes + k-chunk gating of the step-1 GEMM
                from quantum_simulation_b200.parallel import FusedShardedApplyMPO
                fused = FusedShardedApplyMPO(chi, M1.shape[3] * M2.shape[3] * chi, dtype, dev)
                exchange = "fused: P2P push of theta shards (qsb_push_shard) + arrival-gated step-1 GEMM, no NCCL on the data path"
            except Exception as ex:
                if args.exchange == "fused":
                    raise
                exchange += f" (fused path unavailable: {repr(ex)[:120]})"

    def step():
        if fused is not None:
            fused(theta_shard, L, M1, M2, R, out=out)
            return
        if world > 1:      # the exchange step of the sharded matvec: block all-gather of theta
            dist.all_gather_into_tensor(theta, theta_shard)
        op(theta, L, M1, M2, R, out=out)

    # ---- FP64 tensor-pipe peak of this GPU, measured now (no FP64 entry in MEASURED_PEAKS.json)
    peak_tf = _lib.probe_dmma_tflops()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, stage events on the launching stream, clocks sampled meanwhile
    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    for row in evs:
        for e in row:
            e.record()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.launch_counters(reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        _lib.set_stage_events(evs[i])
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.set_stage_events(None)
    sec = e0.elapsed_time(e1) * 1e-3
    launches = _lib.launch_counters()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    ms_step = sec / K * 1e3
    value = flops_total / (sec / K) / 1e9

    g1 = [evs[i][0].elapsed_time(evs[i][1]) for i in range(K)]
    mix = [evs[i][1].elapsed_time(evs[i][2]) for i in range(K)]
    g4 = [evs[i][2].elapsed_time(evs[i][3]) for i in range(K)]
    c = 8.0 if dtype.is_complex else 2.0
    W, d2 = M1.shape[0], M1.shape[3] * M2.shape[3]
    Wl_exec = W - (1 if ident[0] >= 0 else 0)           # GEMM slices actually executed (identity channels skipped)
    Wr_exec = M2.shape[1] - (1 if ident[1] >= 0 else 0)
    g1_flops = c * Wl_exec * rows * chi * d2 * chi      # step-1 GEMM: flops EXECUTED per launch (one launch per step)
    g4_flops = c * chi * (Wr_exec * chi) * (M1.shape[2] * M2.shape[2] * rows)
    g1_ms = sum(g1) / K
    achieved = g1_flops / (g1_ms * 1e-3) / 1e12
    traffic, traffic_src = measured_traffic(args, world)
    roofline = {"bound": "tensor", "kernel": "gemm_tma_kernel (FP64 DMMA), H_eff step 1", "achieved": achieved,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured in this run: register-resident mma.sync.m8n8k4.f64 issue-rate probe "
                               "(qsb_probe_dmma); MEASURED_PEAKS.json has no FP64 entry",
                "algorithmic_flops_per_launch": g1_flops, "avg_launch_ms": g1_ms,
                "channels_executed": {"step1": Wl_exec, "of": W, "step4": Wr_exec, "of_r": int(M2.shape[1])},
                "stage_ms": {"gemm_step1": g1_ms, "mpo_mix": sum(mix) / K, "gemm_step4": sum(g4) / K},
                "step4_tflops": g4_flops / (sum(g4) / K * 1e-3) / 1e12,
                "matvec_frac_of_peak": (flops_exec_local / (ms_step * 1e-3) / 1e12) / peak_tf if peak_tf else None}

    # ---- e2e: the public call with the step's input in pinned HOST memory and the result read back to the host,
    # both copies inside the timed region.  N > 1: every rank uploads ITS shard of theta (the sharded operator's
    # input), the exchange happens on the devices, every rank downloads its rows of the result.
    h_out = torch.empty(n_loc, dtype=dtype).pin_memory()
    if world == 1:
        h_in = torch.empty(n, dtype=dtype).pin_memory()
        h_in.copy_(theta.cpu())
        d_in = torch.empty_like(theta)
    else:
        h_in = torch.empty(n // world, dtype=dtype).pin_memory()
        h_in.copy_(theta_shard.cpu())
        d_in = torch.empty_like(theta_shard)

    e2e_mode = "serial"

    def step_e2e():
        d_in.copy_(h_in, non_blocking=True)
        if fused is not None:
            fused(d_in, L, M1, M2, R, out=out)
        else:
            if world > 1:
                dist.all_gather_into_tensor(theta, d_in)
                op(theta, L, M1, M2, R, out=out)
            else:
                op(d_in, L, M1, M2, R, out=out)
        h_out.copy_(out, non_blocking=True)

    if world == 1:      # the public call itself takes the pinned host vector: upload, matvec, download pipelined
        def step_e2e():                                    # noqa: F811
            op(h_in, L, M1, M2, R, out=h_out)
    elif fused is not None and not args.serial_e2e:
        # N > 1: the fused operator's host-resident call -- shard uploaded in column blocks straight into the gathered
        # buffer, blocks forwarded to the peers as they land, grid-gated step-1 GEMM, row-block download overlapped
        def step_e2e():                                    # noqa: F811
            fused.forward_host(h_in, L, M1, M2, R, h_out)
        e2e_mode = "pipelined"

    for _ in range(2):
        step_e2e()
    if world > 1:
        dist.barrier()
    Ke = max(3, min(K, 10))
    sec_e = time_region(step_e2e, Ke)
    if world > 1:
        t = torch.tensor([sec_e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec_e = float(t.item())
    # how long the HOST needs to enqueue one such step (no device synchronisation inside): if this approaches ms_per_step
    # the e2e number is launch-bound, not copy- or compute-bound
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        step_e2e()
    enqueue_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e = {"value": flops_total / (sec_e / Ke) / 1e9, "unit": UNIT, "h2d_bytes_per_step": n * es,
           "d2h_bytes_per_step": n * es, "ms_per_step": sec_e / Ke * 1e3, "steps": Ke,
           "host_enqueue_ms_per_step": enqueue_ms,
           "bytes_note": "whole-job bytes per step, summed over ranks (each rank moves 1/N of them)",
           "host_pipeline_mode": (next(iter(qsb.ApplyMPO._host_pipes.values()), {}).get("last_mode")
                                  if world == 1 else e2e_mode),
           "call": ("quantum_simulation_b200.ApplyMPO.forward(pinned host psi, out=pinned host) -- chunked upload gating "
                    "the step-1 GEMM, row-block download overlapped" if world == 1 else
                    "quantum_simulation_b200.parallel.FusedShardedApplyMPO.forward_host(pinned host shard, pinned host "
                    "rows) -- column-block upload into the gathered buffer, block-wise P2P forwarding, grid-gated step-1 "
                    "GEMM, row-block download overlapped" if e2e_mode == "pipelined" else
                    "quantum_simulation_b200.parallel.FusedShardedApplyMPO / ShardedApplyMPO, copies around it")
                   + " (torch.ops.qsb200.heff_apply* -> qsb_heff_apply)"}

    # ---- the host link this e2e number rides on, measured now: this rank's share of theta up / of the result down, all
    # ranks at once (they share the host's memory system and PCIe root complexes), alone and both directions together
    try:
        s2 = torch.cuda.Stream(device=dev)
        dd_in, dd_out = torch.empty_like(d_in), torch.empty(n_loc, dtype=dtype, device=dev)

        def both():
            dd_in.copy_(h_in, non_blocking=True)
            s2.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s2):
                h_out.copy_(dd_out, non_blocking=True)
            torch.cuda.current_stream(dev).wait_stream(s2)
        link = {}
        for name, fn in (("h2d", lambda: dd_in.copy_(h_in, non_blocking=True)),
                         ("d2h", lambda: h_out.copy_(dd_out, non_blocking=True)), ("both", both)):
            for _ in range(2):
                fn()
            if world > 1:
                dist.barrier()
            t_ = time_region(fn, 10) / 10
            if world > 1:
                tt = torch.tensor([t_], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_ = float(tt.item())
            link[name + "_ms"] = t_ * 1e3
        link["h2d_GBps_aggregate"] = n * es / (link["h2d_ms"] * 1e-3) / 1e9
        link["d2h_GBps_aggregate"] = n * es / (link["d2h_ms"] * 1e-3) / 1e9
        link["note"] = ("pinned-memory copies of this step's bytes (theta up, result down), every rank its 1/N share "
                        "concurrently, max over ranks; an e2e step cannot be faster than max(device step, 'both')")
        e2e["host_link"] = link
        e2e["floor_ms"] = max(ms_step, link["both_ms"])
        del dd_in, dd_out
    except Exception as ex:
        e2e["host_link"] = {"error": repr(ex)[:200]}

    exec_tf = flops_exec_local * world / (sec / K) / 1e12
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world), "exchange": exchange,
            "clocks": clocks,
            "e2e": e2e, "gpu_launches": launches["total"], "launch_breakdown": launches, "roofline": roofline,
            "tflops": value / 1e3,
            "executed": {"tflops": exec_tf, "flops_per_step": flops_exec_local * world,
                         "dense_flops_per_step": flops_total, "identity_channels": {"L": ident[0], "R": ident[1]},
                         "note": "value counts the dense ncon-order flops the reference executes (effective rate); "
                                 "this is what the kernels execute -- utilisation numbers use it"},
            "frac_of_fp64_tensor_peak": (exec_tf / world) / peak_tf if peak_tf else None}

    if rank == 0 and world == 1 and not args.no_extras:
        line.update(extras(args, qsb, _lib, dev, dtype, L, M1, M2, R, theta, peak_tf))
        if line.get("cublas_dgemm_8192_tflops"):
            line["roofline_frac_vs_cublas_dgemm"] = roofline["achieved"] / line["cublas_dgemm_8192_tflops"]
    if world > 1 and not args.no_extras:
        # every rank takes part (collectives inside); rank 0 keeps the numbers
        del L, R, theta, out, fused
        _lib.release_workspaces()
        torch.cuda.empty_cache()
        if args.model == "heis" and not args.no_sweep:
            line["sweep"] = sharded_sweep_leg(args, qsb, dev, dtype, rank, world)
        if world >= 8 and args.model == "heis" and args.chi == 2048:
            line["large"] = large_configs_leg(args, qsb, _lib, dev, rank, world, local, peak_tf)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sharded_sweep_leg(args, qsb, dev, dtype, rank, world):
    """N > 1, the other half of BASELINE.json's metric: one FULL two-site DMRG sweep of the same config through
    ShardedDMRG (row-sharded environments and Lanczos, fused exchange), timed like the N = 1 'sweep' leg."""
    import contextlib
    import io
    import torch.distributed as dist
    from quantum_simulation_b200.parallel import ShardedDMRG
    try:
        chi, Lsites = args.chi, 100
        key = "heis_f64_L100" if dtype == torch.float64 else "heis_c128_L100"
        mpo = [m.to(dev) for m in load_mpo(key, dtype)]

        def build(L, c):
            with contextlib.redirect_stdout(io.StringIO()):
                m = qsb.MPS(L, 2, c, dev, dtype, seed=1)
            for plist in (m.A, m.B, m.sWeight):                    # identical tensors on every rank (rank 0's)
                for t in plist:
                    dist.broadcast(torch.view_as_real(t.data) if t.is_complex() else t.data, src=0)
            return m

        def sched(c):
            sw = qsb.Sweeps(1)
            sw.set_schedule(maxdim=[c], cutoff=[0.0], krylov_dim=[4], noise=[0.0])
            return sw
        # warm-up on a short chain: NCCL communicators, symmetric-memory machinery, cuSOLVER handles
        small = [mpo[0], mpo[1]] + [mpo[50]] * 12 + [mpo[-2], mpo[-1]]
        ShardedDMRG(build(16, 128), small, device=dev, dtype=dtype)(sched(128), dispon=0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mps = build(Lsites, chi)
        torch.cuda.synchronize()
        t_init = time.perf_counter() - t0
        dm = ShardedDMRG(mps, mpo, device=dev, dtype=dtype)
        dm.profile = False
        dist.barrier()
        torch.manual_seed(1)             # see the N = 1 'sweep' leg: the first Lanczos start comes from the global RNG
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Es, _ = dm(sched(chi), dispon=0, maxit=2)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bd = dm.history["bond_dims"]
        res = {"seconds": float(t.item()), "bond_updates": len(Es), "updates_at_full_chi": int(sum(b == chi for b in bd)),
               "mps_init_on_gpu_s": t_init, "energy_end_of_sweep": float(Es[-1]), "comm": dm.comm,
               "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9,
               "what": f"one full two-site DMRG sweep through ShardedDMRG on {world} GPUs, Heisenberg L={Lsites}, chi={chi}, "
                       f"{args.dtype}, krylov_dim=4, maxit=2, cutoff=0, the same random right-canonical start as the "
                       "N=1 'sweep' leg; wall clock, device synchronise + barrier on both sides, max over ranks"}
        del dm, mps
        qsb.clear_workspaces()
        torch.cuda.empty_cache()
        return res
    except Exception as ex:
        return {"error": repr(ex)[:300]}


def large_configs_leg(args, qsb, _lib, dev, rank, world, local, peak_tf):
    """north_star: "scaling to 8 GPUs at chi >= 4096" -- BASELINE configs 4 (J1-J2 cylinder, MPO bond ~30, float64) and
    5 (Fermi-Hubbard via Jordan-Wigner, d = 4, complex128) at chi = 4096, the sharded matvec with the fused exchange."""
    global _MODEL
    import torch.distributed as dist
    from quantum_simulation_b200.parallel import FusedShardedApplyMPO
    out = {}
    saved = _MODEL
    for model, dts in (("j1j2", "f64"), ("hubbard", "c128")):
        try:
            _MODEL = model
            dtype = torch.float64 if dts == "f64" else torch.complex128
            chi = 4096
            rows = chi // world
            L, M1, M2, R, theta = make_inputs(chi, dtype, dev, rows=rows, row0=rank * rows)
            _lib.tag_identity(L, rank * rows)
            _lib.tag_identity(R)
            ident = (_lib.identity_of(L)[0], _lib.identity_of(R)[0], rank * rows)
            shard = theta.view(world, -1)[rank].clone()
            del theta
            op = FusedShardedApplyMPO(chi, M1.shape[3] * M2.shape[3] * chi, dtype, dev)
            res = torch.empty(rows * M1.shape[2] * M2.shape[2] * chi, dtype=dtype, device=dev)
            for _ in range(2):
                op(shard, L, M1, M2, R, out=res)
            torch.cuda.synchronize()
            sampler = ClockSampler(local)
            if rank == 0:
                sampler.start()
            dist.barrier()
            K = 5
            sec = time_region(lambda: op(shard, L, M1, M2, R, out=res), K)
            dist.barrier()
            clocks = sampler.stop() if rank == 0 else None
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item()) / K
            dense = _lib.heff_flops(torch.empty(0, dtype=dtype, device=dev), L, M1, M2, R) * world
            execd = _lib.heff_flops_executed(torch.empty(0, dtype=dtype, device=dev), L, M1, M2, R, ident) * world
            kf, kc, site, desc = MODELS[model]
            out[model] = {"workload": f"{desc}, chi={chi}, bulk bond {site}|{site + 1}, MPO bonds {M1.shape[0]}-"
                                      f"{M1.shape[1]}-{M2.shape[1]}, d={M1.shape[2]}, {dts}", "n_gpus": world,
                          "ms_per_matvec": sec * 1e3, "tflops_effective": dense / sec / 1e12,
                          "tflops_executed": execd / sec / 1e12,
                          "frac_of_fp64_tensor_peak_per_gpu": execd / sec / 1e12 / world / peak_tf if peak_tf else None,
                          "identity_channels": {"L": ident[0], "R": ident[1]}, "steps": K, "clocks": clocks}
            del L, R, shard, res, op
            _lib.release_workspaces()
            torch.cuda.empty_cache()
        except Exception as ex:
            out[model] = {"error": repr(ex)[:300]}
    _MODEL = saved
    return out


def extras(args, qsb, _lib, dev, dtype, L, M1, M2, R, theta, peak_tf):
    """Library baselines on the same GPU, the other hot-path kernels, and the CPU baseline (rank 0, N = 1)."""
    from oracle import dmrg_oracle as orc
    out = {}
    chi = args.chi
    es = 8 if dtype == torch.float64 else 16
    flops = _lib.heff_flops(theta, L, M1, M2, R)
    # (1) cuBLAS DGEMM 8192^3 and the reference's ncon order executed by torch/cuBLAS on this GPU
    try:
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        s = time_region(lambda: torch.matmul(a, b), 5) / 5
        out["cublas_dgemm_8192_tflops"] = 2 * 8192 ** 3 / s / 1e12
        out["roofline_frac_vs_cublas_dgemm"] = None         # filled by the caller from roofline.achieved
        del a, b
        for _ in range(2):
            orc.apply_mpo(theta, L, M1, M2, R)
        s = time_region(lambda: orc.apply_mpo(theta, L, M1, M2, R), 5) / 5
        out["torch_cublas_same_matvec"] = {"gflops": flops / s / 1e9, "ms": s * 1e3,
                                           "what": "oracle tensordot chain on device='cuda' (cuBLAS + permute copies)"}
        # the reference's contract_backend='auto' resolves to its einsum branch on CUDA (dmrg.py:101-102, 230-248):
        # O(W_l W_r d^2 chi^3) -- what a user of the stock reference gets on this GPU today
        orc.apply_mpo_einsum(theta, L, M1, M2, R)
        s = time_region(lambda: orc.apply_mpo_einsum(theta, L, M1, M2, R), 2) / 2
        out["torch_einsum_auto_backend_same_matvec"] = {"gflops_dense_equivalent": flops / s / 1e9, "ms": s * 1e3,
                                                        "what": "oracle restatement of the reference's einsum branch "
                                                                "on device='cuda'"}
        torch.cuda.empty_cache()
    except Exception as ex:          # library leg is informational only
        out["library_leg_error"] = repr(ex)[:200]
    # (2) Lanczos: one full solve (2 restarts x krylov 4 = 8 matvecs + fused vector ops), vector-op bandwidth
    n = theta.numel()
    op = qsb.ApplyMPO()
    Lh, Rh = hermitian_envs(L, R)
    _lib.tag_identity(Lh), _lib.tag_identity(Rh)
    args_op = (Lh, M1, M2, Rh)
    qsb.lanczos_ground_state(theta, op, args_op, num_restarts=2, krylov_dim=4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, E = qsb.lanczos_ground_state(theta, op, args_op, num_restarts=2, krylov_dim=4)
    torch.cuda.synchronize()
    t_solve = time.perf_counter() - t0
    ws = torch.empty(n, dtype=dtype, device=dev)
    K2 = torch.randn(2, n, dtype=dtype, device=dev)
    scal = torch.zeros(16, 2, dtype=torch.float64, device=dev)
    scratch = torch.zeros(_lib.lz_scratch_bytes(), dtype=torch.uint8, device=dev)
    ops = _lib.ops

    def vec_iter():        # the vector ops of one Krylov iteration: dot, axpy2+norm, scale  (8 n-element passes)
        ops.lz_dot(K2[1], 1, 0, ws, n, scal, 3, 1, scratch)
        ops.lz_axpy_norm(K2, 2, n, ws, n, scal, 2, 4, 0, scratch)
        ops.lz_scale(ws, K2[0], n, scal, 5)
    ws.copy_(theta)
    scal[5, 0] = 1.0
    for _ in range(3):
        vec_iter()
    s = time_region(vec_iter, 20) / 20
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    gbs = 8 * n * es / s / 1e9
    out["lanczos"] = {"solve_ms": t_solve * 1e3, "matvecs": 8, "energy": E,
                      "vector_ops_per_iteration_ms": s * 1e3, "algorithmic_bytes_per_iteration": 8 * n * es,
                      "achieved_gbs": gbs, "hbm_peak_gbs": hbm, "frac_of_hbm_peak": gbs / hbm}
    del ws, K2
    # (3) environment update at the same bond
    A = torch.randn(chi, M1.shape[2], chi, dtype=dtype, device=dev) / chi ** 0.5
    up = qsb.LeftEnvUpdate()
    Lsq = L if L.shape[1] == L.shape[2] else L[:, :, : L.shape[1]]
    for _ in range(2):
        up(Lsq, M1, A)
    s = time_region(lambda: up(Lsq, M1, A), 5) / 5
    ef = _lib.env_flops(Lsq, M1, A, True)
    out["env_update"] = {"ms": s * 1e3, "gflops": ef / s / 1e9, "frac_of_fp64_tensor_peak": ef / s / 1e12 / peak_tf}
    del A
    torch.cuda.empty_cache()
    # (3b) one full bulk bond step of a sweep at this chi: Lanczos (8 matvecs) + env update + theta build + split
    try:
        from quantum_simulation_b200.svd import two_site_split
        d = M1.shape[2]
        Ls, Rs = Lh, Rh
        A = torch.randn(chi, d, chi, dtype=dtype, device=dev) / chi ** 0.5

        def timed(fn, reps=2):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(reps):
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            return best
        t_l = timed(lambda: qsb.lanczos_ground_state(theta, op, (Ls, M1, M2, Rs), num_restarts=2, krylov_dim=4))
        t_e = timed(lambda: up(Ls, M1, A))
        t_t = timed(lambda: torch.matmul(A.reshape(chi * d, chi), A.reshape(chi, d * chi)))
        # the split is timed on a block with a DMRG-like Schmidt spectrum (exponential decay: negligible weight beyond
        # chi, what every bond of a converged sweep looks like -> range-compressed path) and, separately, on the
        # full-rank random theta of this bench (an unconverged first half-sweep -> the compression is rejected and the
        # full Gram/eigh path answers)
        from quantum_simulation_b200 import svd as qsvd
        m2 = chi * d
        gsp = torch.Generator(device=dev).manual_seed(11)
        a_ = torch.randn(m2, m2, dtype=dtype, device=dev, generator=gsp)
        U0 = torch.linalg.qr(a_)[0]
        V0 = torch.linalg.qr(a_.T.contiguous())[0]
        sv = torch.exp(-torch.arange(m2, dtype=torch.float64, device=dev) * (240.0 / m2))
        th_dmrg = (U0 * sv.to(dtype)) @ V0.conj().T
        th_dmrg = th_dmrg / torch.linalg.norm(th_dmrg)
        del a_, U0, V0
        qsvd.clear_caches()
        before = dict(qsvd.stats)
        adapt = qsvd.SKETCH_ADAPT
        qsvd.SKETCH_ADAPT = False            # full-width compression: what a bond with numerical rank ~ chi costs
        t_s = timed(lambda: two_site_split(th_dmrg, chi, 0.0))
        took_sketch = qsvd.stats["sketch_accepted"] > before["sketch_accepted"]
        qsvd.SKETCH_ADAPT = adapt            # rank-adaptive width: this synthetic block has numerical rank ~300
        t_s_adapt = timed(lambda: two_site_split(th_dmrg, chi, 0.0), reps=3)
        th = theta.view(chi * d, d * chi)
        t_s_full = timed(lambda: two_site_split(th, chi, 0.0, sketch=False))
        del th_dmrg
        step = t_l + t_e + t_t + t_s
        out["bond_step"] = {"lanczos_8_matvecs_ms": t_l * 1e3, "env_update_ms": t_e * 1e3, "theta_build_ms": t_t * 1e3,
                            "two_site_split_ms": t_s * 1e3, "two_site_split_path": "range-compressed Gram/eigh" if took_sketch else "full Gram/eigh",
                            "two_site_split_full_gram_ms": t_s_full * 1e3,
                            "two_site_split_rank_adaptive_ms": t_s_adapt * 1e3, "total_ms": step * 1e3,
                            "sweep_estimate_s": 2 * (100 - 1) * step,
                            "note": "bulk bond of Heisenberg L=100 at full chi; a sweep is 198 such steps (edge bonds "
                                    "are cheaper), so sweep_estimate_s is an upper estimate of the sweep time"}
        del A, Ls, Rs
        torch.cuda.empty_cache()
    except Exception as ex:
        out["bond_step"] = {"error": repr(ex)[:200]}
    # (3c) the stock-builder flavour of the same config: complex128 MPO (Sx, Sy, Sz), same chi
    try:
        if dtype == torch.float64:
            Lc, M1c, M2c, Rc, thc = make_inputs(chi, torch.complex128, dev)
            _lib.tag_identity(Lc), _lib.tag_identity(Rc)
            outc = torch.empty_like(thc)
            for _ in range(2):
                op(thc, Lc, M1c, M2c, Rc, out=outc)
            s = time_region(lambda: op(thc, Lc, M1c, M2c, Rc, out=outc), 5) / 5
            fc = _lib.heff_flops(thc, Lc, M1c, M2c, Rc)
            fe = _lib.heff_flops_executed(thc, Lc, M1c, M2c, Rc, (_lib.identity_of(Lc)[0], _lib.identity_of(Rc)[0], 0))
            out["complex128_same_config"] = {"gflops": fc / s / 1e9, "ms": s * 1e3, "executed_gflops": fe / s / 1e9,
                                             "frac_of_fp64_tensor_peak": fe / s / 1e12 / peak_tf,
                                             "note": "Heisenberg1DMPOBuilder-style complex MPO; 8 flop per complex MAC"}
            del Lc, Rc, thc, outc
            torch.cuda.empty_cache()
    except Exception as ex:
        out["complex128_same_config"] = {"error": repr(ex)[:200]}
    # (3d) the other half of BASELINE.json's metric: one FULL two-site DMRG sweep, measured (Heisenberg config only)
    if args.model == "heis" and not args.no_sweep:
        try:
            import contextlib
            import io
            Lsites = 100
            del L, R, theta, Lh, Rh, args_op
            _lib.release_workspaces()
            qsb.clear_workspaces()
            torch.cuda.empty_cache()
            with contextlib.redirect_stdout(io.StringIO()):
                t0 = time.perf_counter()
                mps = qsb.MPS(Lsites, 2, chi, dev, dtype, seed=1)          # canonicalised on the GPU
                torch.cuda.synchronize()
                t_init = time.perf_counter() - t0
                sw = qsb.Sweeps(1)
                sw.set_schedule(maxdim=[chi], cutoff=[0.0], krylov_dim=[4], noise=[0.0])
                key = "heis_f64_L100" if dtype == torch.float64 else "heis_c128_L100"
                dm = qsb.DMRG(mps, [m.to(dev) for m in load_mpo(key, dtype)], device=dev, dtype=dtype)
                # At L=100, chi=2048 the reference's MPS constructor (and its mirror) yields a ZERO centre tensor (the
                # carry's squared norm overflows), so the first Lanczos start vector is re-seeded from the GLOBAL RNG
                # (lanczos_solver.py:150-153): seed it, or two runs differ by 2e-6 (r01's drift; tools/first_bond_diag.py)
                torch.manual_seed(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                Es, _ = dm(sw, dispon=0, maxit=2)
                torch.cuda.synchronize()
                t_sweep = time.perf_counter() - t0
            bd = dm.history["bond_dims"]
            out["sweep"] = {"seconds": t_sweep, "bond_updates": len(Es), "updates_at_full_chi": int(sum(b == chi for b in bd)),
                            "mps_init_on_gpu_s": t_init, "energy_end_of_sweep": float(Es[-1]),
                            "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9,
                            "rng": "torch.manual_seed(1) before the sweep (the zero centre tensor of the L=100/chi=2048 "
                                   "start makes the first Lanczos start a global-RNG draw, as in the reference)",
                            "what": f"one full two-site DMRG sweep, Heisenberg L={Lsites}, chi={chi}, {args.dtype}, "
                                    "krylov_dim=4, maxit=2, cutoff=0, random right-canonical start; wall clock with a "
                                    "device synchronise on both sides"}
            del dm, mps
            qsb.clear_workspaces()
            torch.cuda.empty_cache()
        except Exception as ex:
            out["sweep"] = {"error": repr(ex)[:200]}
    # (4) CPU baseline: the oracle on the host cores, bounded sample of the same workload
    cores = os.cpu_count() or 1
    try:
        rows = chi if chi <= 1024 else max(256, chi // 4)
        ts, fl, threads, kind = cpu_matvec_time(chi, dtype, 3, rows=rows)
        sec = min(ts[1:])
        out["cpu_baseline"] = {"value": fl / sec / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": f"rows [0,{rows}) of {chi} of the same chi={chi} matvec, best of 2 after 1 "
                                         f"warm-up, " + ("the unmodified reference's ApplyMPO('ncon') from baseline/_ref"
                                                         if kind == "reference" else "torch-CPU oracle (ncon order)")
                                         + f", {threads} threads",
                               "ms_sample": sec * 1e3}
    except Exception as ex:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": repr(ex)[:200]}
    return out


def main():
    global _MODEL
    global _DENSE_ENV
    args = parse()
    _MODEL = args.model
    _DENSE_ENV = args.dense_env
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
