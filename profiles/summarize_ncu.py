#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the small summaries committed under profiles/:
  <out>.metrics.json   selected raw metrics per captured launch
  <out>.hot.txt        the hottest SASS instructions with executed counts and top stall reasons
Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_v1
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second",
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    raw = [r for r in raw if len(r) > 10]
    hdr, units, launches = raw[0], raw[1], raw[2:]
    summary = []
    for row in launches:
        d = {"kernel": row[hdr.index("Kernel Name")]}
        for k in KEYS + [h for h in hdr if h.startswith("smsp__warp_issue_stalled") and h.endswith("per_warp_active.pct")]:
            if k in hdr:
                v = row[hdr.index(k)]
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    pass
                d[k] = [v, units[hdr.index(k)]]
        summary.append(d)
    json.dump(summary, open(out + ".metrics.json", "w"), indent=1)

    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hdr = src[1]
    body = []
    for r in src[2:]:
        if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    iS, iE, iT, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[iE]) for r in body)
    samples = sum(int(r[iN]) for r in body)
    agg = {}
    with open(out + ".hot.txt", "w") as f:
        f.write(f"# {src[0][1]}\n# {len(body)} SASS instructions, {total} warp-instructions executed, {samples} stall samples\n")
        f.write("# idx | SASS | executed (M) | avg threads | samples | top stalls\n")
        for k, r in enumerate(body):
            for i in stalls:
                if r[i] not in ("", "0"):
                    agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
            if int(r[iE]) > total * 0.004:
                st = sorted(((hdr[i][6:], int(r[i])) for i in stalls if r[i] not in ("", "0")), key=lambda x: -x[1])[:3]
                f.write(f"{k:5d} | {r[iS].strip()[:64]:64s} | {int(r[iE]) / 1e6:8.2f} | {r[iT]:>4s} | {r[iN]:>6s} | {st}\n")
        f.write("# stall samples by reason: " + json.dumps(dict(sorted(agg.items(), key=lambda x: -x[1]))) + "\n")
    print("wrote", out + ".metrics.json", out + ".hot.txt")


if __name__ == "__main__":
    main()
