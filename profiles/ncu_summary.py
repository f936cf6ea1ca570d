"""Summarises an ncu report for profiles/: key raw metrics + SASS opcode mix + hottest instructions.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_kernel.txt"""
import csv
import collections
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name"), " id", d.get("ID"))
    for k in KEYS:
        if k in d:
            print("  %-86s %s %s" % (k, d[k], units[hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
iS, iN, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
ops_n, ops_s = collections.Counter(), collections.Counter()
tot_n = tot_s = 0
insts = []
for r in rows[2:]:
    if len(r) <= iN or not r[iN].isdigit():
        continue
    op = r[iSrc].strip().split()[0]
    if op.startswith("@"):
        op = r[iSrc].strip().split()[1]
    op = op.split(".")[0]
    n, s = int(r[iN]), int(r[iS])
    ops_n[op] += n; ops_s[op] += s; tot_n += n; tot_s += s
    insts.append((s, n, r[0][-5:], r[iSrc].strip()))
print("\n== SASS opcode mix (warp instructions executed, share of stall samples)")
for op, n in ops_n.most_common(18):
    print("  %-10s %6.2f%% of instructions   %6.2f%% of samples" % (op, 100.0 * n / tot_n, 100.0 * ops_s[op] / max(1, tot_s)))
print("\n== hottest 25 instructions by samples (samples, executed, addr, sass)")
for s, n, a, t in sorted(insts, reverse=True)[:25]:
    print("  %7d %12d  %s  %s" % (s, n, a, t[:90]))
