"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the markdown table kept under profiles/.

    python tools/summarise_launches.py gpurun_out/launches.csv "title" > profiles/rNN_launches_*.md
"""
import collections
import csv
import re
import sys


def main(path, title):
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        val = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ms = val * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = re.sub(r'^void ', '', name)
        name = re.sub(r'^b200::', '', name)
        rows.append((name, ms))
    tot = sum(ms for _, ms in rows)
    agg = collections.OrderedDict()
    for name, ms in rows:
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
    print(f'# {title}')
    print(f'{len(rows)} consecutive launches, {tot:.1f} ms of device time. Cold-cache and serialised (ncu replays each '
          f'kernel): compare shares, not absolutes.\n')
    print('| ms | share | launches | kernel |\n|---|---|---|---|')
    for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'| {ms:.2f} | {100 * ms / tot:.1f}% | {n} | `{name}` |')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'ncu launch list')
