#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics (raw page) + instruction/stall distribution by SASS segment (source page).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [segment_size]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; seg = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_bytes.sum', 'launch__grid_size', 'launch__block_size', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct', 'smsp__warp_issue_stalled_imc_miss_per_warp_active.pct']
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if h in keys: print("  %-75s %-12s %s" % (h, units[i], r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(r for r in rows if "Address" in r)
body = [r for r in rows if len(r) == len(h) and r[0].startswith("0x")]
I, T, S = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[I]) for r in body) or 1; tots = sum(int(r[S]) for r in body) or 1
print("SASS instructions %d, warp-insts executed %d, avg threads/inst %.2f" % (len(body), tot, sum(int(r[T]) for r in body) / tot))
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
for i in range(0, len(body), seg):
    b = body[i:i + seg]
    ie = sum(int(r[I]) for r in b); te = sum(int(r[T]) for r in b); ss = sum(int(r[S]) for r in b)
    if ie * 200 < tot and ss * 200 < tots: continue
    ops = {}
    for r in b:
        w = r[1].split(); op = w[1] if w[0].startswith('@') else w[0]
        ops[op] = ops.get(op, 0) + int(r[I])
    st = {c: sum(int(r[h.index(c)] or 0) for r in b) for c in stall_cols}
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]; tst = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5d inst%% %5.1f thr/inst %5.1f stall%% %5.1f | %s | %s" % (i, 100 * ie / tot, te / max(ie, 1), 100 * ss / tots,
          " ".join("%s:%.0f%%" % (k, 100 * v / max(ie, 1)) for k, v in top), " ".join("%s:%.0f%%" % (k[6:], 100 * v / max(ss, 1)) for k, v in tst)))
