#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full --import-source on) into a text summary for profiles/:
per kernel the roofline-relevant raw metrics and the instructions with the most stall samples.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_blocks", "launch__occupancy_limit_warps",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = [f"# summary of {rep} (ncu --set full --clock-control none --import-source on)", ""]
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        lines.append("== " + d.get("Kernel Name", "?"))
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:78s} {d[k]:>16s} {units[hdr.index(k)]}")
        lines.append("")
    src = page(rep, "source")
    kern, table = None, []
    blocks = []
    for r in src:
        if r and r[0] == "Kernel Name":
            if kern:
                blocks.append((kern, table))
            kern, table = r[1], []
        elif r and r[0].startswith("0x"):
            table.append(r)
    if kern:
        blocks.append((kern, table))
    seen = set()
    for kern, table in blocks:
        if kern in seen:
            continue
        seen.add(kern)
        tot = sum(int(t[2] or 0) for t in table) or 1
        lines.append(f"== top stall-sample instructions: {kern} (total samples {tot}, {len(table)} SASS instr)")
        for t in sorted(table, key=lambda t: -int(t[2] or 0))[:22]:
            lines.append(f"  {100.0 * int(t[2] or 0) / tot:5.1f}%  exec={t[5]:>11s}  {t[1].strip()}")
        lines.append("")
    open(dst, "w").write("\n".join(lines))
    print("\n".join(lines[:6]), f"\n... wrote {dst}")


if __name__ == "__main__":
    main()
