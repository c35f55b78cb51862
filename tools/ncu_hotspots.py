#!/usr/bin/env python3
"""Top stall hot spots from `ncu -i rep --page source --csv` (SASS view) for the first kernel in the report."""
import csv
import subprocess
import sys


def main(rep, topn=40, kernel_index=0):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'rows': [], 'hdr': None}
            blocks.append(cur)
        elif cur is not None and row and row[0] == 'Address':
            cur['hdr'] = row
        elif cur is not None and cur['hdr'] and row:
            cur['rows'].append(row)
    b = blocks[kernel_index]
    h = {n: i for i, n in enumerate(b['hdr'])}
    tot = sum(int(r[h['# Samples']]) for r in b['rows'])
    print(b['name'][:90], 'total samples', tot, 'instructions', len(b['rows']))
    # cumulative by region between barriers
    reg, acc, inst = 0, 0, 0
    for r in b['rows']:
        acc += int(r[h['# Samples']]); inst += 1
        if 'BAR.SYNC' in r[h['Source']]:
            print('  region %d: %5.1f%% of samples, %d instr' % (reg, 100.0 * acc / tot, inst)); reg += 1; acc = 0; inst = 0
    print('  region %d: %5.1f%% of samples, %d instr' % (reg, 100.0 * acc / tot, inst))
    idx = sorted(range(len(b['rows'])), key=lambda i: -int(b['rows'][i][h['# Samples']]))[:topn]
    for i in sorted(idx):
        r = b['rows'][i]
        print('%5d %5.2f%%  %s' % (i, 100.0 * int(r[h['# Samples']]) / tot, r[h['Source']].strip()[:100]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
