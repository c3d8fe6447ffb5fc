"""Aggregates warp-stall samples of one kernel from an .ncu-rep (source page) by reason and by top SASS lines."""
import collections
import csv
import subprocess
import sys


def main(path, kernel_regex, top=14):
  out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel_regex],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
  ia, isamp = hdr.index('Source'), hdr.index('# Samples')
  stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
  tot = collections.Counter()
  lines = []
  for r in rows:
    if len(r) <= isamp or not r[isamp].isdigit():
      continue
    st = {h: int(r[hdr.index(h)] or 0) for h in stalls}
    for k, v in st.items():
      tot[k] += v
    lines.append((int(r[isamp]), r[ia], st))
  allsum = sum(tot.values())
  print('stall reasons:', ', '.join(f'{k[6:]} {v / allsum * 100:.1f}%' for k, v in tot.most_common(8)))
  for n, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
    print(f'{n:6d} {src[:80]:80s}', dict(sorted(((k[6:], v) for k, v in st.items() if v), key=lambda kv: -kv[1])[:2]))


if __name__ == '__main__':
  main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 14)
