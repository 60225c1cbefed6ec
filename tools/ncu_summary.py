"""Summarise an ncu capture: key raw metrics + per-segment SASS instruction/stall shares.
usage: ncu_summary.py raw.csv sass.csv [out_prefix]"""
import csv
import sys

raw, sass = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
print('kernel:', d[hdr.index('Kernel Name')][:100])
for k in keys:
    if k in hdr:
        print(f'  {k:75s} {d[hdr.index(k)]} {units[hdr.index(k)]}')
rows = list(csv.reader(open(sass)))
hdr = rows[1]
iex, ismp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and '(Not Issued)' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ('Kernel Name', 'Address'):
        if r and r[0] == 'Kernel Name' and data:
            break
        continue
    data.append(r)
tot = sum(int(r[iex]) for r in data)
tots = sum(int(r[ismp]) for r in data)
print(f'total warp-instructions {tot}, samples {tots}, sass lines {len(data)}')
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print('stall samples:', ', '.join(f'{k[6:]}={v / tots * 100:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
segs = []
for n, r in enumerate(data):
    ex, sm = int(r[iex]), int(r[ismp])
    if segs and 0.7 < (ex + 1) / (segs[-1]['ex0'] + 1) < 1.4:
        s = segs[-1]
        s['n'] += 1; s['ex'] += ex; s['smp'] += sm; s['end'] = n
    else:
        segs.append({'start': n, 'end': n, 'ex0': ex, 'n': 1, 'ex': ex, 'smp': sm})
for s in segs:
    if s['ex'] / tot > 0.005:
        top = sorted(data[s['start']:s['end'] + 1], key=lambda r: -int(r[ismp]))[0]
        print(f"sass {s['start']:4d}-{s['end']:4d} n={s['n']:4d} exec/instr={s['ex0']:>11d} inst={s['ex'] / tot * 100:5.1f}% "
              f"samples={s['smp'] / tots * 100:5.1f}%  hottest: {top[isrc][:60]}")
