"""profiles/summarize.py -- turn ncu outputs brought back in gpurun_out/ into the small CSVs committed here.

    python profiles/summarize.py launches gpurun_out/X_launches.csv profiles/Y_launches.csv
    python profiles/summarize.py full gpurun_out/X_prof.ncu-rep profiles/Y_full.csv
"""
import collections
import csv
import subprocess
import sys

FULL = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, data = None, []
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        a = agg.setdefault(d['Kernel Name'][:70], [0, 0.0])
        a[0] += 1
        a[1] += float(d['Metric Value'].replace(',', ''))
    unit = data[0]['Metric Unit'] if data else ''
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', f'total_{unit}', f'avg_{unit}', 'share'])
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, c, round(t, 1), round(t / c, 1), round(t / tot, 4)])
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(w) for w in FULL if w in hdr]
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    for r in rows[2:]:
        print({hdr[i]: r[i][:40] for i in idx})


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
