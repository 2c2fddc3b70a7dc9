"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches / total ms / share.

usage: python tools/launch_summary.py launches.csv [last_n_launches]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = re.sub(r'^void |\(anonymous namespace\)::', '', name)
        rows.append((name, ns))
    if len(sys.argv) > 2:
        rows = rows[-int(sys.argv[2]):]
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f'{len(rows)} launches, {tot / 1e6:.1f} ms')
    print('| kernel | launches | ms | share |\n|---|---:|---:|---:|')
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f'| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |')


if __name__ == '__main__':
    main()
