#!/usr/bin/env python
"""ncu report -> short text summary of the metrics the roofline discussion uses (run here: no GPU needed).
    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__waves_per_multiprocessor", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100], "| grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("   %-90s %s %s" % (k, r[i], units[i]))
        # pipe utilisation and the warp-stall breakdown (what an issue-bound kernel is waiting for)
        for i, k in enumerate(hdr):
            if k in KEYS:
                continue
            if ("inst_executed_pipe_" in k and k.endswith("pct_of_peak_sustained_active")) or \
                    ("issue_stalled" in k and k.endswith("per_warp_active.pct")) or k in (
                        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
                        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"):
                try:
                    if float(r[i].replace(",", "")) == 0.0:
                        continue
                except ValueError:
                    pass
                print("   %-90s %s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1])
