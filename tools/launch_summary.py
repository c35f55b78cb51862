#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: average duration and share per kernel."""
import collections
import csv
import re
import sys


def main(path, steps=1):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
        agg.setdefault((name, row['Grid Size'], row['Block Size']), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print('%-28s grid=%-16s blk=%-12s n=%-3d avg=%8.1f us  share=%5.1f%%' % (k[0][:28], k[1], k[2], len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print('total %.1f us over %d launches' % (tot, sum(len(v) for v in agg.values())))


if __name__ == '__main__':
    main(sys.argv[1])
