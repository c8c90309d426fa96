"""Summarises `ncu --set full` reports (gpurun_out/*.ncu-rep) into small tracked files under profiles/:
   <name>_raw.csv (selected raw metrics), <name>_opmix.txt (SASS opcode mix + top stall lines) and traffic.json
   (DRAM bytes per launch of each kernel, used by bench.py for roofline.traffic).
Usage: python profiles/extract_ncu.py <report.ncu-rep> <name> <kernel-key> <images-in-capture> [kernel-regex]
(kernel-regex selects one kernel of a report that holds several: passed to `ncu -i ... -k regex:<kernel-regex>`)"""
import collections, csv, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = ('gpu__time_duration', 'dram__bytes', 'gpu__dram_throughput', 'sm__throughput', 'sm__warps_active', 'launch__', 'smsp__inst_executed.sum',
        'smsp__issue_active', 'l1tex__t_sectors_pipe_lsu_mem_global', 'l1tex__data_bank_conflicts', 'lts__throughput', 'l1tex__throughput',
        'smsp__average_warps_issue_stalled', 'sm__pipe_alu_cycles_active', 'sm__pipe_fma_cycles_active', 'sm__inst_executed_pipe_lsu')


def main():
    rep, name, key, images = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    sel = ['-k', 'regex:' + sys.argv[5]] if len(sys.argv) > 5 else []
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    with open(os.path.join(HERE, name + '_raw.csv'), 'w') as f:
        for h in hdr:
            if h.startswith(KEEP) and 'not_issued' not in h:
                f.write(f'{h},{u[h]},{m[h]}\n')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h2 = rows[1]; si, ii, st = h2.index('Source'), h2.index('Instructions Executed'), h2.index('# Samples')
    nxt = [i for i, r in enumerate(rows) if i > 1 and r and r[0] == 'Kernel Name']     # a filtered multi-kernel report repeats the block
    if nxt: rows = rows[:nxt[0]]
    ops, tot = collections.Counter(), 0
    for r in rows[2:]:
        if len(r) <= ii: continue
        parts = r[si].split()
        if not parts: continue
        op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
        n = int(r[ii] or 0); ops[op.split('.')[0]] += n; tot += n
    with open(os.path.join(HERE, name + '_opmix.txt'), 'w') as f:
        f.write(f'kernel {rows[0][1] if len(rows[0]) > 1 else key}\ntotal warp instructions {tot}\n')
        for op, n in ops.most_common(24): f.write(f'{op:12s} {n:12d} {100 * n / max(tot, 1):5.1f}%\n')
        f.write('\ntop stall-sample lines (# samples, executed, SASS)\n')
        for r in sorted(rows[2:], key=lambda r: -int(r[st] or 0))[:15]: f.write(f'{r[st]:>6} {r[ii]:>9} {r[si].strip()[:100]}\n')
    def mb(k):
        v = float(m[k]); un = u[k].lower()
        return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[un]
    tpath = os.path.join(HERE, 'traffic.json')
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    t[key] = {'dram_bytes_per_launch': mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum'), 'images_in_capture': images,
              'duration_us_under_ncu': float(m['gpu__time_duration.sum']), 'report': os.path.basename(rep), 'source': name + '_raw.csv'}
    json.dump(t, open(tpath, 'w'), indent=1)
    print(name, 'dram MB', t[key]['dram_bytes_per_launch'] / 1e6, 'instr', tot)


if __name__ == '__main__':
    main()
