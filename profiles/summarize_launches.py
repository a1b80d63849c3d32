#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, mean
duration and share of the total (shares are what is comparable: ncu serialises and runs cold)."""
import collections
import csv
import sys


def main(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v if u in ('ns', 'nsecond') else v * 1e3 if u in ('us', 'usecond') else v * 1e6 if u in ('ms', 'msecond') else v
        agg.setdefault(row['Kernel Name'], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f'# {path}: {sum(len(v) for v in agg.values())} launches, {tot / 1e3:.1f} us total (serialised, cold)')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'{100 * sum(v) / tot:5.1f}%  n={len(v):4d}  avg={sum(v) / len(v) / 1e3:9.1f} us  {k[:110]}')


if __name__ == '__main__':
    main(sys.argv[1])
