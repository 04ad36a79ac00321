"""Per-kernel summary of an ncu launch list: time share, launches, DRAM bytes and achieved DRAM GB/s of every kernel
of one bench step (conv, rulebook, voxelize scatter, clustering, kNN ... -- the `north_star` evidence list).

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-plot --no-train --no-cluster
  python tools/summarise_launches.py gpurun_out/launches.csv [peak_GBs] > profiles/rNN_ncu_launches_summary.txt

Times under ncu are cold-cache and serialised: compare SHARES with bench.py's live numbers, not absolutes.  The GB/s
column is dram bytes / duration of the same replayed launch, i.e. the HBM rate that launch really drew."""
import collections
import csv
import json
import os
import re
import sys

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    mp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    peak = json.load(open(mp))['hbm_gbs'] if os.path.exists(mp) else 7700.0
rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 10]
ix = {h: i for i, h in enumerate(rows[0])}
per_launch = collections.OrderedDict()
for r in rows[1:]:
    if r[ix['ID']] == 'ID':
        continue
    d = per_launch.setdefault((r[ix['Process ID']], r[ix['ID']]), {'kernel': r[ix['Kernel Name']]})
    v = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'usecond': 1, 'nsecond': 1e-3,
             'ms': 1e3, 'msecond': 1e3}.get(unit, 1)
    d[r[ix['Metric Name']]] = v * scale


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name[:110]


agg = collections.OrderedDict()
for l in per_launch.values():
    a = agg.setdefault(short(l['kernel']), [0.0, 0, 0.0, 0.0])
    a[0] += l.get('gpu__time_duration.sum', 0.0)
    a[1] += 1
    a[2] += l.get('dram__bytes_read.sum', 0.0)
    a[3] += l.get('dram__bytes_write.sum', 0.0)
total = sum(a[0] for a in agg.values())
print(f'# {len(per_launch)} launches, {total / 1e3:.3f} ms (ncu: cold-cache, serialised); HBM peak used for the % column: {peak:.1f} GB/s')
print(f'# {"share":>6} {"total_us":>10} {"launches":>8} {"dram_rd_MB":>10} {"dram_wr_MB":>10} {"GB/s":>8} {"%peak":>6}  kernel')
for k, (us, n, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0.0
    print(f'  {100 * us / total:5.2f}% {us:10.1f} {n:8d} {rd / 1e6:10.1f} {wr / 1e6:10.1f} {gbs:8.1f} {100 * gbs / peak:5.1f}%  {k}')
