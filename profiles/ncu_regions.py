"""Dev helper: key pipe/stall metrics + region-level instruction distribution of one `ncu --set full` report.
Usage: python profiles/ncu_regions.py <report.ncu-rep> [top-lines]"""
import csv, re, subprocess, sys

rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
pat = (r'gpu__time_duration.sum|smsp__inst_executed.sum$|sm__issue_active.avg.pct|pipe_alu_cycles_active.avg.pct_of_peak_sustained_active|'
       r'pipe_fma_cycles_active.avg.pct_of_peak_sustained_active|pipe_lsu.avg.pct_of_peak_sustained_active|average_warps_issue_stalled.*_per_issue_active|'
       r'warps_active.avg.pct|bank_conflicts_pipe_lsu_mem_shared(_op_ld|_op_st)?.sum|dram__bytes_(read|write).sum$|launch__registers|launch__occupancy_limit')
for n, u, v in zip(rows[0], rows[1], rows[2]):
    if re.search(pat, n) and 'not_issued' not in n:
        print(f'{n} [{u}] {v}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]; si, ii, st = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
out = [(int(r[ii] or 0), int(r[st] or 0), r[si].strip()) for r in rows[2:] if len(r) > ii]
tot = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print('total warp instr', tot, 'samples', ts)
prev = None; start = 0; acc = cnt = smp = 0
for i, (n, s, _) in enumerate(out + [(-1, 0, '')]):
    if prev is None or abs(n - prev) > 0.02 * max(prev, 1):
        if prev is not None and acc > 0.004 * tot:
            print(f'{start:5d}-{i - 1:5d} n={prev:9d} instrs={cnt:4d} {100 * acc / tot:5.1f}% of instr, {100 * smp / max(ts, 1):5.1f}% of samples')
        start = i; acc = cnt = smp = 0
    prev = n; acc += max(n, 0); cnt += 1; smp += s
print('top stall lines')
for i in sorted(range(len(out)), key=lambda i: -out[i][1])[:top]:
    print(f'{i:5d} {out[i][1]:6d} {out[i][0]:9d} {out[i][2][:90]}')
