"""Sum the DRAM traffic of the conv launches of ONE forward from an ncu csv and write profiles/r02_conv_traffic.json
(keyed by workload, mode and a hash of the conv kernel sources: bench.py reports `traffic` only while the hash matches).

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:k_conv --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_layers.py cfg2_2M f16
  python tools/summarise_conv_traffic.py gpurun_out/conv_traffic.csv cfg2_2M f16 4      # 4 forwards in that run
"""
import collections
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treelearn_b200._lib import conv_source_hash   # noqa: E402

path, workload, mode, n_fwd = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per_launch = collections.OrderedDict()
for r in rows[1:]:
    d = per_launch.setdefault(r[ix['ID']], {'kernel': r[ix['Kernel Name']]})
    v = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(unit, 1)
    d[r[ix['Metric Name']]] = v * scale
launches = list(per_launch.values())
per_fwd = len(launches) // n_fwd
last = launches[-per_fwd:]                      # the last (warm) forward
rd = sum(l.get('dram__bytes_read.sum', 0) for l in last)
wr = sum(l.get('dram__bytes_write.sum', 0) for l in last)
us = sum(l.get('gpu__time_duration.sum', 0) for l in last)
out = {'workload': workload, 'mode': mode, 'conv_source_sha1': conv_source_hash(), 'launches_per_step': per_fwd, 'dram_bytes_read_per_step': int(rd),
       'dram_bytes_write_per_step': int(wr), 'dram_bytes_per_step': int(rd + wr), 'ncu_kernel_us_per_step': round(us, 1),
       'how': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none '
              '-k regex:k_conv on tools/profile_layers.py; last forward of the run'}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r02_conv_traffic.json')
caps = json.load(open(dst)) if os.path.exists(dst) else []
caps = [c for c in (caps if isinstance(caps, list) else [caps]) if (c.get('workload'), c.get('mode')) != (workload, mode)] + [out]
json.dump(caps, open(dst, 'w'), indent=1)
print(json.dumps(out))
