"""Summarise an .ncu-rep (raw + source pages) into a small text report: python scripts/ncu_summary.py rep [out.txt]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
def run(page):
    return subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(run('raw'))))
hdr, units, vals = raw[0], raw[1], raw[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__cycles_active.avg']
print(f'# {rep}', file=out)
for k in keys:
    for i, h in enumerate(hdr):
        if h == k:
            print(f'{k} = {vals[i]} {units[i]}', file=out)
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
        try:
            if float(vals[i]) >= 0.05:
                print(f'{h} = {vals[i]}', file=out)
        except ValueError:
            pass
src = list(csv.reader(io.StringIO(run('source'))))
h2 = src[1]; data = src[2:]
ix = {h: i for i, h in enumerate(h2)}
tot = sum(float(r[ix['Instructions Executed']] or 0) for r in data)
samp = sum(float(r[ix['# Samples']] or 0) for r in data)
byop = collections.Counter(); wf = collections.Counter()
for r in data:
    toks = r[ix['Source']].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    byop[op.split('.')[0]] += float(r[ix['Instructions Executed']] or 0)
    if op.startswith(('LDS', 'STS')):
        wf[op] += float(r[ix['L1 Wavefronts Shared']] or 0)
print(f'static SASS instructions = {len(data)} ({len(data) * 16} B); executed warp instructions = {tot:.4g}; samples = {samp:.0f}', file=out)
print('instruction mix: ' + ', '.join(f'{o} {c / tot * 100:.1f}%' for o, c in byop.most_common(14)), file=out)
print('shared wavefronts: ' + ', '.join(f'{o} {c:.3g}' for o, c in wf.most_common(8)) + f'; total {sum(wf.values()):.4g}', file=out)
for k in ['stall_barrier', 'stall_branch_resolving', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_no_inst', 'stall_not_selected',
          'stall_selected', 'stall_short_sb', 'stall_wait', 'stall_lg', 'stall_dispatch', 'stall_sleep', 'stall_membar']:
    if k in ix:
        print(f'{k} = {sum(float(r[ix[k]] or 0) for r in data) / samp * 100:.1f}% of samples', file=out)
