"""Summarise an .ncu-rep: headline metrics, stall reasons, SASS hot regions. Usage: ncu_summary.py file.ncu-rep [kernel-index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'sm__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__waves_per_multiprocessor',
        'smsp__cycles_active.avg', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
for r in rows[2:]:
    print("-" * 100)
    for w in want:
        if w in hdr:
            print(f"{w:75s} {r[hdr.index(w)][:80]}  [{rows[1][hdr.index(w)]}]")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = [r for r in rows[2:] if len(r) == len(h)]
iex = h.index('Instructions Executed'); isamp = h.index('# Samples'); ith = h.index('Avg. Threads Executed')
tot = sum(int(r[iex] or 0) for r in data); ts = sum(int(r[isamp] or 0) for r in data)
print(f"SASS: {len(data)} instr, executed {tot}, samples {ts}")
for name in ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_no_inst', 'stall_math', 'stall_lg',
             'stall_barrier', 'stall_membar', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_mio', 'stall_sleep']:
    if name in h:
        v = sum(int(r[h.index(name)] or 0) for r in data)
        if v: print(f"   {name:26s} {v:7d}  {100.0*v/max(1,ts):5.1f}%")
for b in range(0, len(data), 100):
    e = sum(int(r[iex] or 0) for r in data[b:b+100]); s = sum(int(r[isamp] or 0) for r in data[b:b+100])
    th = [float(r[ith]) for r in data[b:b+100] if r[ith] and int(r[iex] or 0) > 0]
    if e: print(f"   sass[{b:4d}..] inst {e:10d} ({100.0*e/tot:4.1f}%) samples {s:6d} ({100.0*s/max(1,ts):4.1f}%) avg-threads {sum(th)/max(1,len(th)):.1f}")
