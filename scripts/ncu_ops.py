"""Aggregates an `ncu --page source --csv` dump by SASS opcode (executed warp instructions, stall samples)."""
import collections
import csv
import sys


def main(path, warps):
  rows = list(csv.reader(open(path)))
  hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
  ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
  ops, samp, tot = collections.Counter(), collections.Counter(), 0
  for r in rows:
    if len(r) <= ie or not r[ie].isdigit():
      continue
    parts = r[ia].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    n = int(r[ie])
    ops[op] += n
    tot += n
    samp[op] += int(r[isamp] or 0)
  print('total', tot, 'per warp', tot / warps)
  for op, n in ops.most_common(30):
    print(f'{op:10s} {n:10d} {n / tot * 100:5.1f}%  per-warp {n / warps:7.1f} samples {samp[op]}')


if __name__ == '__main__':
  main(sys.argv[1], float(sys.argv[2]))
