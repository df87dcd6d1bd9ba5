"""Summarise an ncu report: key raw metrics + the hottest SASS instructions grouped by source line.
   python tools/ncu_hot.py gpurun_out/prof_x.ncu-rep [topN]"""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor']
for w in want:
    if w in H:
        print('%-78s %s %s' % (w, V[H.index(w)], U[H.index(w)]))
for h in H:
    if 'issue_stalled' in h and 'per_issue_active' in h or 'tensor' in h and 'pct' in h:
        try:
            if float(V[H.index(h)]) > 0.3: print('%-78s %s' % (h, V[H.index(h)]))
        except ValueError: pass
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if hi:
    H = rows[hi[0]]; body = rows[hi[0] + 1:]
    si, ii, ss = H.index('Source'), H.index('Instructions Executed'), H.index('Warp Stall Sampling (All Samples)')
    tot_i = sum(float(r[ii] or 0) for r in body if len(r) > ii); tot_s = sum(float(r[ss] or 0) for r in body if len(r) > ss)
    agg = collections.Counter(); aggs = collections.Counter()
    for r in body:
        if len(r) <= ss: continue
        op = re.sub(r'\s+', ' ', r[si]).strip().split(' ')
        op = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
        agg[op] += float(r[ii] or 0); aggs[op] += float(r[ss] or 0)
    print('--- opcode mix (share of executed warp instructions | share of stall samples), total inst %.3g' % tot_i)
    for op, c in agg.most_common(topn):
        print('  %-28s %5.1f%%  | %5.1f%%' % (op, 100 * c / max(tot_i, 1), 100 * aggs[op] / max(tot_s, 1)))
