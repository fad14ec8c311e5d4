"""Summaries of ncu output for profiles/ (run here, no GPU needed).

  python scripts/ncu_summarize.py full  gpurun_out/prof.ncu-rep  profiles/rN_ncu_full_kernels.csv
  python scripts/ncu_summarize.py list  gpurun_out/launches.csv  profiles/rN_launches_summary.md "<command>"
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def short(name):
    n = name.split('(')[0]
    for pre in ('void sk::', 'sk::', 'void '):
        if n.startswith(pre):
            n = n[len(pre):]
    return n.replace('(anonymous namespace)::', '')


def full(rep, out):
    if rep.endswith('.csv'):      # already exported on the GPU box: ncu -i X.ncu-rep --page raw --csv > X.csv
        raw = open(rep).read()
    else:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, key, scale=1.0):
        i = col.get(key)
        if i is None or r[i] in ('', 'n/a'):
            return ''
        v = float(r[i].replace(',', ''))
        u = units[i]
        if key.startswith('dram__bytes'):
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            return v * mult / 1e6
        if key == 'gpu__time_duration.sum':
            mult = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(u, 1)
            return v * mult
        return v * scale
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['id', 'kernel', 'grid', 'block', 'regs', 'dyn_smem', 'duration_us', 'dram_read_MB', 'dram_write_MB',
                    'dram_pct', 'tensor_pct', 'sm_pct', 'warps_active_pct', 'dram_GBps'])
        for r in data:
            dur = get(r, 'gpu__time_duration.sum')
            rd, wr = get(r, 'dram__bytes_read.sum'), get(r, 'dram__bytes_write.sum')
            gbps = (rd + wr) / dur * 1e3 if dur and rd != '' else ''   # MB/us = TB/s -> GB/s
            grid = 'x'.join(x.strip() for x in r[col['Grid Size']].strip('()').split(',') if x.strip() != '1') or '1' if 'Grid Size' in col else ''
            block = r[col['Block Size']].strip('()').split(',')[0].strip() if 'Block Size' in col else ''
            w.writerow([r[col['ID']], short(r[col['Kernel Name']]), grid, block,
                        get(r, 'launch__registers_per_thread'), get(r, 'launch__shared_mem_per_block_dynamic'),
                        round(dur, 3) if dur != '' else '', rd, wr,
                        get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                        get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') or
                        get(r, 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active'),
                        get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
                        get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
                        round(gbps, 1) if gbps != '' else ''])
    print('wrote', out, len(data), 'kernels')


def launches(path, out, command):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(io.StringIO(''.join(lines))))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        u = r.get('Metric Unit', 'ns')
        us = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1e-3)
        k = short(r['Kernel Name']).split('<')[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    n = sum(a[0] for a in agg.values())
    with open(out, 'w') as f:
        f.write(f"# ncu launch list -- `{command}`\n\n{n} consecutive launches from the steady state.  Times are cold-cache and "
                f"serialised (ncu replays each kernel): compare SHARES, not absolutes.\n\n"
                f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {us / 1e3:.3f} | {100 * us / total:.1f} % |\n")
        f.write(f"| **total** | {n} | {total / 1e3:.3f} | 100 % |\n")
    print('wrote', out)


if __name__ == '__main__':
    if sys.argv[1] == 'full':
        full(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
