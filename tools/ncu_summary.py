#!/usr/bin/env python3
"""Summarise an `ncu --set full` capture (run here, no GPU needed): per captured launch the duration, DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum), tensor/FP64 pipe activity, registers, L2 hit rate -- written as a
markdown table and, for the dominant kernel, as the `traffic` entry bench.py reports in its roofline object
(profiles/r02_ncu_traffic.json; bench.py never invents this number: no capture, traffic = null).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --key heis/chi2048/f64/n1/canonical \
        --kernel gemm_tma_kernel --md profiles/r02_gemm_tma_ncu_full.md
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "launch__shared_mem_per_block_dynamic"]


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    v = num(v)
    if v is None:
        return None
    unit = (unit or "").strip().lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "": 1}.get(unit, 1)


def to_ms(v, unit):
    v = num(v)
    if v is None:
        return None
    unit = (unit or "").strip().lower()
    return v * {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--key", default=None, help="bench configuration key for profiles/r02_ncu_traffic.json")
    ap.add_argument("--kernel", default="gemm_tma_kernel")
    ap.add_argument("--md", default=None)
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head = rows[0]
    units = rows[1]
    body = rows[2:]
    col = {n: i for i, n in enumerate(head)}
    kname = col.get("Kernel Name")
    out = []
    for r in body:
        if len(r) != len(head):
            continue
        e = {"kernel": r[kname]}
        for m in WANT:
            if m in col:
                e[m] = (r[col[m]], units[col[m]])
        out.append(e)
    lines = ["| # | kernel | ms | DRAM read MB | DRAM write MB | regs | grid x block | DMMA sub-pipe active % of elapsed | L2 hit % | DRAM % of peak |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    picked = []
    for i, e in enumerate(out):
        ms = to_ms(*e.get("gpu__time_duration.sum", (None, None)))
        rd = to_bytes(*e.get("dram__bytes_read.sum", (None, None)))
        wr = to_bytes(*e.get("dram__bytes_write.sum", (None, None)))
        pipe = None
        a = e.get("SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg")
        b = e.get("sm__cycles_elapsed.avg")
        if a and b and num(a[0]) and num(b[0]):                 # DMMA sub-pipe active cycles / elapsed cycles
            pipe = round(100.0 * num(a[0]) / num(b[0]), 2)
        short = e["kernel"].split("(")[0][-70:]
        lines.append(f"| {i} | `{short}` | {ms:.4f} | {(rd or 0) / 1e6:.1f} | {(wr or 0) / 1e6:.1f} | "
                     f"{e.get('launch__registers_per_thread', ('', ''))[0]} | {e.get('launch__grid_size', ('', ''))[0]} x "
                     f"{e.get('launch__block_size', ('', ''))[0]} | {pipe if pipe is not None else ''} | "
                     f"{e.get('lts__t_sector_hit_rate.pct', ('', ''))[0]} | "
                     f"{e.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', ('', ''))[0]} |")
        if args.kernel in e["kernel"] and rd is not None and wr is not None:
            picked.append({"ms": ms, "dram_bytes": rd + wr, "kernel": short})
    text = "\n".join(lines)
    print(text)
    if args.md:
        with open(args.md, "w") as f:
            f.write(f"# ncu --set full --clock-control none: {os.path.basename(args.rep)}\n\n{args.note}\n\n{text}\n\n"
                    "All metric names available in the capture for the first launch of the dominant kernel:\n\n")
            if picked:
                first = next(e for e in out if args.kernel in e["kernel"])
                for m in WANT:
                    if m in first:
                        f.write(f"* `{m}` = {first[m][0]} {first[m][1]}\n")
    if args.key and picked:
        path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        z = json.load(open(path)) if os.path.exists(path) else {}
        first = picked[0]                       # the step-1 GEMM is the first launch of the dominant kernel in a step
        z[args.key] = {"dram_bytes": first["dram_bytes"], "ms_under_ncu": first["ms"], "kernel": first["kernel"],
                       "all_launches": picked, "source": f"ncu --set full capture {os.path.basename(args.rep)} (this round)"}
        json.dump(z, open(path, "w"), indent=1)
        print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
