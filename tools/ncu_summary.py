"""Summarise an `ncu --set full` report into the JSON kept under profiles/:
   ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv && python tools/ncu_summary.py /tmp/raw.csv > profiles/...json
One object per profiled launch with the metrics the roofline discussion uses (duration, DRAM bytes, DRAM / SM / L1
throughput, tensor-pipe activity, registers, achieved occupancy, issue activity)."""
import csv
import json
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_subpipe_tc_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size']

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
out = []
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    o = {'Kernel Name': r[col['Kernel Name']]}
    for k in ('Grid Size', 'Block Size'):
        if k in col:
            o[k] = r[col[k]]
    for w in WANT:
        if w in col and r[col[w]] != '':
            o[w] = ('%s %s' % (r[col[w]], units[col[w]])).strip()
    # tensor-pipe activity: ONLY the recipe's metric (B200_PROFILING.md: sm__pipe_tensor_cycles_active, % of peak, plain
    # counter).  The round-1 summary also swept every column with "tensor" in its name, among them the sampled
    # TriageCompute `..._realtime` variant, which reads 12.5 and 49.7 for two identical launches of one kernel (the
    # 46.8 / 12.5 pair the review found): not a usable measurement, no longer reported.
    for n in names:
        if n.startswith('sm__pipe_tensor_cycles_active') and 'pct_of_peak' in n and n.split('.')[1] == 'avg' and r[col[n]] != '':
            o[n] = ('%s %s' % (r[col[n]], units[col[n]])).strip()
    out.append(o)
json.dump(out, sys.stdout, indent=1)
