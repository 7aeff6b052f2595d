#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`): headline metrics, stall reasons, hottest source lines.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv, io, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[2:]

def main():
    rep = sys.argv[1]
    h, rows = raw(rep)
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    for row in rows:
        print("=" * 100)
        for k in keys:
            if k in h:
                print(f"{k:75s} {row[h.index(k)]}")
        ps = [(float(row[i].replace(",", "")), n) for i, n in enumerate(h)
              if "smsp__pcsamp_warps_issue_stalled" in n and not n.endswith("_not_issued") and row[i] not in ("", "n/a")]
        tot = sum(v for v, _ in ps) or 1
        print("-- warp state samples (all samples)")
        for v, n in sorted(ps, reverse=True)[:10]:
            print(f"   {v / tot * 100:5.1f}%  {n.replace('smsp__pcsamp_warps_issue_stalled_', '')}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(io.StringIO(out)))
        hh = r[0]
        si = hh.index("Source") if "Source" in hh else 1
        cand = [c for c in hh if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Samples" in c]
        col = hh.index(cand[0]) if cand else None
        ei = hh.index("Instructions Executed") if "Instructions Executed" in hh else None
        print("source columns:", hh[:12])
        if col is not None:
            body = [x for x in r[1:] if len(x) > col and x[col].replace(",", "").replace(".", "").isdigit()]
            tots = sum(float(x[col].replace(",", "")) for x in body) or 1
            body.sort(key=lambda x: -float(x[col].replace(",", "")))
            for x in body[:n]:
                print(f"{float(x[col].replace(',', '')) / tots * 100:5.1f}%  exec={x[ei] if ei is not None else ''}  {x[si][:110]}")

if __name__ == "__main__":
    main()
