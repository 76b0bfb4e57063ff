#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed) and an ncu launch list into the text kept under profiles/.

usage: python scripts/ncu_summary.py gpurun_out/c2_full.ncu-rep [gpurun_out/launches_c2.csv] > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu summary of `{rep}`\n")
    print("`ncu --set full --clock-control none --import-source on` (per-launch, cold cache, serialised; times are not bench values)\n")
    for r in data:
        print(f"## {r[hdr.index('Kernel Name')]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for w in WANT:
            if w in hdr:
                print(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("\nwarp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + "\n")
    if len(sys.argv) > 2:
        print("## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n")
        agg = {}
        with open(sys.argv[2]) as f:
            lines = [l for l in f if l.startswith('"')]
        rd = list(csv.reader(lines))
        h = rd[0]
        kn, mv = h.index("Kernel Name"), h.index("Metric Value")
        for r in rd[1:]:
            name = r[kn].split("(")[0][-70:]
            t = float(r[mv].replace(",", ""))
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += t
        tot = sum(a[1] for a in agg.values())
        print("| kernel | launches | total ns | share |\n|---|---:|---:|---:|")
        for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
            print(f"| {name} | {n} | {t:.0f} | {t / tot * 100:.1f}% |")


if __name__ == "__main__":
    main()
