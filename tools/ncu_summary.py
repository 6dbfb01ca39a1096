#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + opcode mix / stall samples from the SASS page.
usage: python scripts_ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'inst_executed', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:3]:
    print('KERNEL', r[hdr.index('Kernel Name')][:160])
    for k in KEYS:
        if k in hdr:
            print('  %-75s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
    for i, k in enumerate(hdr):
        if re.search(r'smsp__average_warps?_issue_stalled_.*_per_issue_active|smsp__average_warp_latency_issue_stalled', k):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.15:
                print('  stall %-68s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''), v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hidx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hidx[0]]
end = hidx[1] - 1 if len(hidx) > 1 else len(rows)
ci = {k: i for i, k in enumerate(h)}
ops, samp = collections.Counter(), collections.Counter()
tot = tots = 0
body = [r for r in rows[hidx[0] + 1:end] if len(r) >= len(h)]
for r in body:
    s = r[ci['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s)
    op = m.group(2) if m else s
    op = '.'.join(op.split('.')[:2]) if op.startswith(('MUFU', 'LDS', 'STS', 'LDG', 'STG')) else op.split('.')[0]
    ie = int(r[ci['Instructions Executed']] or 0)
    sm = int(r[ci['Warp Stall Sampling (All Samples)']] or 0)
    ops[op] += ie; samp[op] += sm; tot += ie; tots += sm
print('static SASS instrs %d, executed warp-instrs %d, samples %d' % (len(body), tot, tots))
for op, c in ops.most_common(28):
    print('  %-12s %10d (%4.1f%%)  samples %5.1f%%' % (op, c, 100.0 * c / tot, 100.0 * samp[op] / max(1, tots)))
stall_cols = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
agg = collections.Counter()
for r in body:
    for k in stall_cols:
        try:
            agg[k] += int(r[ci[k]] or 0)
        except ValueError:
            pass
print('stall reasons (all samples):', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / max(1, sum(agg.values()))) for k, v in agg.most_common(9)))
