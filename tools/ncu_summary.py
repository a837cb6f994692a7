#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed): tools/ncu_summary.py <rep> <out.txt> [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__mio_inst_issued.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "sm__cycles_elapsed.avg",
    "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "local_load", "local_store",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep} (extracted by tools/ncu_summary.py)\n")
        for r in rows[2:]:
            d = dict(zip(hdr, zip(units, r)))
            f.write(f"\n## kernel: {d.get('Kernel Name', ('', '?'))[1]}  grid {d.get('Grid Size', ('', '?'))[1]} block {d.get('Block Size', ('', '?'))[1]}\n")
            for k in hdr:
                if any(k == w or (w in k and len(w) < 14) for w in KEYS):
                    f.write(f"{k} [{d[k][0]}] = {d[k][1]}\n")
            f.write("-- warp stall reasons (cycles per issued instruction)\n")
            for k in hdr:
                m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
                if m and float(d[k][1] or 0) >= 0.01:
                    f.write(f"stall {m.group(1)} = {d[k][1]}\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
