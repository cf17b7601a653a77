"""Turns an `ncu --set full` report into the text summaries committed under profiles/.
usage: summarize.py <report.ncu-rep> <out prefix>"""
import csv
import subprocess
import sys

rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__icc_requests.sum", "sm__icc_requests_lookup_hit.sum", "sm__icc_requests_lookup_miss.sum",
    "gcc__requests.sum", "smsp__thread_inst_executed.sum", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]
with open(prefix + "_raw_summary.txt", "w") as f:
    f.write(f"# selected raw metrics of {rep} (ncu --set full --clock-control none), one block per captured launch\n")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        f.write(f"\n== {r[idx['Kernel Name']]}  grid {r[idx['Grid Size']] if 'Grid Size' in idx else ''}\n")
        for w in WANT:
            if w in idx:
                f.write(f"{w:85s} {r[idx[w]]:>18s} {units[idx[w]]}\n")
print("wrote", prefix + "_raw_summary.txt")
