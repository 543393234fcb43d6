"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md cites.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more-substrings...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "sm__cycles_elapsed.avg ",
        "smsp__cycles_active.avg", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct", "sm__issue_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_uniform", "sm__inst_executed_pipe_xu",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "dram__bytes_write.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_write.sum ", "lts__t_bytes.sum.per_second", "smsp__inst_executed.sum ",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__warps_active.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency", "sm__throughput.avg.pct", "gpu__compute_memory_throughput",
        "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "smsp__inst_issued.avg.per_cycle_active"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("==", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in hdr:
            kk = k + " "
            if any(s in kk for s in KEYS + extra):
                v = d[k]
                if v not in ("", "0", "0.000000"):
                    print(f"  {k} = {v} {u[k]}")


if __name__ == "__main__":
    main()
