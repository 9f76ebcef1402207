#!/usr/bin/env python
"""Text summary of one kernel of an .ncu-rep capture: the counters quoted in DESIGN.md / profiles/README.md and the
source lines with the largest share of executed warp instructions (needs -lineinfo and --import-source on).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--title "..."] [--top 18] > profiles/rNN_<what>_summary.txt
"""
import argparse
import csv
import subprocess

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_red.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--title", default="")
    ap.add_argument("--top", type=int, default=18)
    a = ap.parse_args()
    rows = ncu_csv(a.report, "raw")
    units, vals = dict(zip(rows[0], rows[1])), dict(zip(rows[0], rows[2]))
    if a.title:
        print(a.title)
    print("kernel: " + vals.get("Kernel Name", "?"))
    for k in KEYS:
        if k in vals:
            print("%-85s %-16s %s" % (k, units.get(k, ""), vals[k]))
    for h, v in vals.items():
        if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(v) > 0.3:
                    print("%-85s %-16s %s" % (h, "inst", v))
            except ValueError:
                pass
    src = ncu_csv(a.report, "source", ["--print-source", "cuda,sass"])[3:]
    agg, cur = {}, None

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in src:
        if r and r[0]:
            cur = (r[0], ",".join(r[1:]).split(",-,-,")[0].strip()[:100])
        elif len(r) > 8 and r[2].startswith("0x"):
            e = agg.setdefault(cur, [0, 0])
            e[0] += num(r[7])
            e[1] += num(r[8])
    total = sum(e[0] for e in agg.values())
    if total:
        print("\nsource lines by share of executed warp instructions (share, avg active threads, line, text):")
        for (ln, code), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
            if n:
                print("%5.1f%%  thr=%4.1f  L%-4s %s" % (100.0 * n / total, t / n, ln, code))


if __name__ == "__main__":
    main()
