#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per launch, instruction mix, stall reasons, hottest SASS lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_hot]"""
import collections, csv, subprocess, sys, io

rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ldgsts.sum']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} = {r[i]} {units[i]}")
    for i, k in enumerate(hdr):   # everything about the shared-memory / LSU pipes
        if ('shared' in k or 'pipe_lsu' in k or 'inst_executed_pipe_uniform' in k or 'l1tex__lsu_writeback' in k) and k not in keys:
            print(f"{k} = {r[i]} {units[i]}")
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
byop, tot, stall, lines, cls, clsn = collections.Counter(), 0, collections.Counter(), [], collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ie])
    except ValueError:
        continue
    toks = r[ia].split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    byop[op] += n
    tot += n
    cls[n] += n
    clsn[n] += 1
    st = {}
    for i in cols:
        try:
            v = int(r[i])
        except ValueError:
            v = 0
        if v:
            stall[hdr[i]] += v
            st[hdr[i][6:]] = v
    try:
        lines.append((int(r[isamp]), n, r[ia][:100], st))
    except ValueError:
        pass
print("total warp instructions (first launch):", tot)
print("mix:", ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in byop.most_common(18)))
print("by execution count:", "; ".join(f"{n}x{clsn[n]} ({100 * v / tot:.0f}%)" for n, v in sorted(cls.items(), key=lambda kv: -kv[1])[:8]))
S = sum(stall.values())
print("stalls:", ", ".join(f"{k[6:]} {100 * v / S:.1f}%" for k, v in stall.most_common(10)))
lines.sort(key=lambda x: -x[0])
for smp, n, s, st in lines[:nhot]:
    print(f"{smp:5d} {n:8d}  {s}  {st}")
