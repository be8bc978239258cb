#!/usr/bin/env python
"""Turns the ncu artefacts brought back in gpurun_out/ into the text summaries kept under profiles/:
  python profiles/summarize.py launches gpurun_out/launches_rX.csv  > profiles/launches_rX.md
  python profiles/summarize.py kernel   gpurun_out/prof_*.ncu-rep   > profiles/<kernel>_rX.txt
"""
import collections
import csv
import subprocess
import sys

SETUP = ("k_db_", "k_synth", "k_seg_fill", "k_seg_count", "k_reads_sample", "k_read_offsets", "k_tab_", "k_grp_", "k_fill",
         "at::", "k_gt_")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    step = {k: v for k, v in agg.items() if not any(s in k for s in SETUP)}
    tot = sum(a[1] for a in step.values())
    print(f"# ncu launch list: {path}\n")
    print(f"{len(data)} launches captured (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: "
          "compare shares, not absolutes).  Setup kernels (synthetic data, DB build) are listed separately.\n")
    print("## hot-path kernels (share of the summed hot-path kernel time)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(step.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:90]}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")
    print("\n## setup kernels (outside the timed region)\n")
    print("| kernel | launches | total us |\n|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if k not in step:
            print(f"| `{k[:90]}` | {n} | {t:.1f} |")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__cycles_elapsed.max"]


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    ki = H.index("Kernel Name")
    print(f"ncu --set full --clock-control none: {path}")
    for r in rows[2:]:
        print(f"\n== {r[ki]} ==")
        for w in WANT:
            if w in H:
                print(f"  {w:62s} {r[H.index(w)]:>16s} {rows[1][H.index(w)]}")
        print("  stall reasons (warps per issue-active cycle):")
        for i, h in enumerate(H):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.1:
                    print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:6.2f}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
