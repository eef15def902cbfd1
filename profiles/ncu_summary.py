"""Key numbers of every kernel in an ncu report, as markdown.
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "branch uniformity %"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency / issued instruction (cycles)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}
print(f"# ncu summary of `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`)\n")
for r in rows[2:]:
    print(f"## {r[col['Kernel Name']].split('(')[0]}  (launch id {r[col['ID']]})\n")
    print("| metric | value |\n|---|---|")
    for k, name in KEYS:
        if k in col and r[col[k]] != "":
            print(f"| {name} (`{k}`) | {r[col[k]]} {u[col[k]]} |")
    print()
