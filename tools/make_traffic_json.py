"""profiles/traffic.json from the raw ncu exports of the three hot kernels (run_profiles.sh): DRAM bytes per launch at
16384^2, stamped with the sha256 of the kernels' source so that bench.py quotes it only for the binary it was captured on.

    python tools/make_traffic_json.py <k_stepNx3_full.csv> <k_step2x_full.csv> <k_step_pair_full.csv> > profiles/traffic.json
"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def metrics(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            x = float(v.replace(',', ''))
        except ValueError:
            continue
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}.get(u, None)
        out[h] = x * scale if scale and h.startswith('dram__bytes') else x
        if h == 'gpu__time_duration.sum':
            out['duration'] = f'{v} {u}'
    return out


names = ['k_stepNx3', 'k_step2x', 'k_step_pair']
keys = {'k_stepNx3': 'k_stepNx3_dram_bytes_per_launch_16384', 'k_step2x': 'k_step2x_dram_bytes_per_launch_16384',
        'k_step_pair': 'dram_bytes_per_launch_16384'}
steps = {'k_stepNx3': 3, 'k_step2x': 2, 'k_step_pair': 1}
res = {'lattice': '16384x16384', 'source_sha256': bench.source_sha(),
       'algorithmic_bytes_per_step': 16384 * 16384 * 144}
for name, path in zip(names, sys.argv[1:4]):
    m = metrics(path)
    rd, wr = m['dram__bytes_read.sum'], m['dram__bytes_write.sum']
    res[keys[name]] = rd + wr
    res[keys[name] + '_source'] = (f'{os.path.basename(path)} ({name}, ncu --set full --clock-control none): dram__bytes_read.sum '
                                   f'{rd / 1e9:.3f} GB + dram__bytes_write.sum {wr / 1e9:.3f} GB per launch of {steps[name]} time '
                                   f'step(s) = {(rd + wr) / steps[name] / 16384 / 16384:.1f} B per cell update; '
                                   f'gpu__time_duration {m.get("duration", "?")}')
print(json.dumps(res, indent=1))
