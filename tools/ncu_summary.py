#!/usr/bin/env python3
"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of
metrics the roofline discussion uses. usage: tools/ncu_summary.py rep.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]

BYTE_UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out.append(f"== {name}  (launch id {r[hdr.index('ID')]})")
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = r[i]
                out.append(f"  {k:86s} {r[i]:>18s} {units[i]}")
        try:
            t = float(vals["gpu__time_duration.sum"])
            # ncu picks a unit per COLUMN value range (Mbyte for the reads, Gbyte for the writes
            # of the same launch): convert before adding
            rd = float(vals["dram__bytes_read.sum"]) * BYTE_UNITS[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(vals["dram__bytes_write.sum"]) * BYTE_UNITS[units[hdr.index("dram__bytes_write.sum")]]
            out.append(f"  -> dram traffic per launch = {(rd + wr) / 1e6:.4g} Mbyte"
                       f" over {t:.4g} {units[hdr.index('gpu__time_duration.sum')]}")
            out.append(f"  -> warp execution efficiency = "
                       f"{float(vals['smsp__thread_inst_executed_per_inst_executed.ratio']) / 32 * 100:.1f} %")
        except (KeyError, ValueError):
            pass
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
