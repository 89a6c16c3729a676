"""Summarise an .ncu-rep (raw + source pages) into a short text report: python scripts/ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__cycles_elapsed.avg', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_active.avg']
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    print(f"== kernel: {name[:100]}", file=out)
    for i, h in enumerate(hdr):
        if h in want or ('issue_stalled' in h and h.endswith('_per_issue_active.ratio')):
            try:
                if 'issue_stalled' in h and float(vals[i]) < 0.05:
                    continue
            except ValueError:
                pass
            print(f"  {h} [{units[i]}] = {vals[i]}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = [r for r in csv.reader(src.splitlines()) if len(r) >= 6 and r[0].startswith('0x')]
tot = sum(int(r[2]) for r in srows) or 1
c = Counter(); ex = Counter()
for r in srows:
    toks = r[1].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    c[op] += int(r[2]); ex[op] += int(r[5])
print(f"-- warp stall samples by opcode (total {tot}); warp-level instructions executed", file=out)
for op, v in c.most_common(12):
    print(f"  {op:10s} {100 * v / tot:5.1f}%   executed {ex[op]}", file=out)
print("-- hottest instructions", file=out)
for r in sorted(srows, key=lambda r: -int(r[2]))[:12]:
    print(f"  {int(r[2]):6d} samples  exec {r[5]:>10s}  {r[1].strip()[:80]}", file=out)
