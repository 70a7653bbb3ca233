"""Summarise an ncu report (read here on the CPU box): python tools/ncu_summary.py REP.ncu-rep OUT.md [title]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_issued.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none), one row per captured launch", ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append(f"## {d.get('Kernel Name', '?')}  (id {d.get('ID', '?')})")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {u[k]} |")
        lines.append("")
        lines.append("warp stall reasons (warps per issue-active cycle, > 0.15):")
        lines.append("")
        for k in hdr:
            if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if v > 0.15:
                    lines.append(f"- {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
