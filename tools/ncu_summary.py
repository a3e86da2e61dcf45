"""Extract the metrics this project reports from an .ncu-rep (read here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__stack_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for li, val in enumerate(rows[2:]):
        name = val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## launch {li}: {name[:90]}")
        for h, u, v in zip(hdr, units, val):
            if h in KEYS or any(h.startswith(k) and h.endswith(".ratio") for k in ("smsp__average_warp",)) \
                    or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main(sys.argv[1])
