"""Digest an `ncu --page raw --csv` export into a per-launch table (JSON + markdown) for profiles/.

    python tools/ncu_summary.py gpurun_out/r01_step_fp16x3_raw.csv profiles/r01_ncu_step_fp16x3

Times are ncu's (cold cache, serialised, clocks as found): shares and traffic are meaningful, absolutes are not a
bench number.
"""
import csv
import json
import sys

COLS = {
    'name': 'Kernel Name',
    'grid': 'launch__grid_size',
    'block': 'launch__block_size',
    'regs': 'launch__registers_per_thread',
    'ms': 'gpu__time_duration.sum',
    'dram_rd': 'dram__bytes_read.sum',
    'dram_wr': 'dram__bytes_write.sum',
    'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'tensor_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'tensor_pct_rt': 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l2_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm_mhz': 'sm__cycles_elapsed.avg.per_second',
    'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
}
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12,
         'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3,
         'Ghz': 1e3, 'Mhz': 1.0, 'GHz': 1e3, 'MHz': 1.0, 'hz': 1e-6}


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return None


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    find = {}
    for key, want in COLS.items():
        idx = [i for i, h in enumerate(hdr) if h == want or h.endswith('.' + want)]
        find[key] = idx[0] if idx else None
    out = []
    for r in body:
        rec = {}
        for key, i in find.items():
            if i is None:
                rec[key] = None
                continue
            if key == 'name':
                rec[key] = r[i].split('(')[0].replace('void ', '').replace('(anonymous namespace)::', '')
                continue
            v = num(r[i])
            if v is not None and units[i] in SCALE and key in ('ms', 'dram_rd', 'dram_wr', 'sm_mhz'):
                v *= SCALE[units[i]]
            rec[key] = v
        if rec['dram_rd'] is not None and rec['dram_wr'] is not None and rec['ms']:
            rec['traffic_bytes'] = rec['dram_rd'] + rec['dram_wr']
            rec['dram_gbs'] = rec['traffic_bytes'] / (rec['ms'] * 1e-3) / 1e9
        out.append(rec)
    total = sum(r['ms'] or 0 for r in out)
    by_kernel = {}
    for r in out:
        k = by_kernel.setdefault(r['name'], {'launches': 0, 'ms': 0.0, 'traffic_bytes': 0.0, 'tensor_ms': 0.0})
        k['launches'] += 1
        k['ms'] += r['ms'] or 0
        k['traffic_bytes'] += r.get('traffic_bytes') or 0
        k['tensor_ms'] += (r['ms'] or 0) * (r['tensor_pct'] if r['tensor_pct'] is not None else (r['tensor_pct_rt'] or 0)) / 100.0
    for k in by_kernel.values():
        k['share'] = k['ms'] / total if total else None
        k['traffic_bytes_per_launch'] = k['traffic_bytes'] / k['launches']
        k['tensor_pipe_pct_time_weighted'] = 100.0 * k['tensor_ms'] / k['ms'] if k['ms'] else None
        del k['tensor_ms']
    weighted = sum((r['ms'] or 0) * (r['tensor_pct'] or 0) for r in out) / total if total else None
    json.dump({'source': src, 'total_ms': total, 'tensor_pipe_pct_time_weighted_whole_step': weighted, 'by_kernel': by_kernel,
               'launches': out}, open(dst + '.json', 'w'), indent=1)
    with open(dst + '.md', 'w') as f:
        f.write(f'ncu --set full --clock-control none, one device-resident step ({src}); total {total:.2f} ms under ncu\n\n')
        f.write('tensor pipe % = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active (= hmma sub-pipe cycles / 4 / active cycles; '
                'the `_realtime` variant of the metric, which round 1 tabulated, disagrees with it by up to 4x between launches of '
                'identical work and is not used)\n\n')
        f.write('| # | kernel | grid | regs | ms | share | DRAM rd MB | DRAM wr MB | DRAM GB/s | tensor pipe % | SM % | L2 % |\n')
        f.write('|---|---|---|---|---|---|---|---|---|---|---|---|\n')
        for i, r in enumerate(out):
            def fm(v, p=1):
                return '-' if v is None else f'{v:.{p}f}'
            f.write(f"| {i} | {r['name']} | {fm(r['grid'], 0)} | {fm(r['regs'], 0)} | {fm(r['ms'], 3)} | "
                    f"{fm(100 * (r['ms'] or 0) / total)}% | {fm((r['dram_rd'] or 0) / 1e6)} | {fm((r['dram_wr'] or 0) / 1e6)} | "
                    f"{fm(r.get('dram_gbs'), 0)} | {fm(r['tensor_pct'] if r['tensor_pct'] is not None else r['tensor_pct_rt'])} | "
                    f"{fm(r['sm_pct'])} | {fm(r['l2_pct'])} |\n")
        f.write('\n| kernel | launches | ms | share | traffic/launch MB | tensor pipe % (time-weighted) |\n|---|---|---|---|---|---|\n')
        for n, k in by_kernel.items():
            f.write(f"| {n} | {k['launches']} | {k['ms']:.3f} | {100 * k['share']:.1f}% | {k['traffic_bytes_per_launch'] / 1e6:.1f} | "
                    f"{k['tensor_pipe_pct_time_weighted']:.1f} |\n")
    with open(dst + '.md', 'a') as f:
        f.write(f'\nTensor pipe, time-weighted over the whole step: {weighted:.1f} %\n')
    print(open(dst + '.md').read())


if __name__ == '__main__':
    main()
