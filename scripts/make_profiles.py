"""Turns the ncu outputs of `scripts/gpu_check.sh <tag>` (gpurun_out/) into the committed
summaries under profiles/: launch list, per-kernel headline metrics, per-launch DRAM traffic."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw_rows(rep):
  out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  return hdr, units, rows[2:]


def to_bytes(v, unit):
  return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main(tag, rnd):
  g = os.path.join(ROOT, 'gpurun_out')
  p = os.path.join(ROOT, 'profiles')
  os.makedirs(p, exist_ok=True)
  shutil.copy(os.path.join(g, f'launches_{tag}.csv'), os.path.join(p, f'{rnd}_launches_ncu.csv'))
  rep = os.path.join(g, f'prof_{tag}.ncu-rep')
  with open(os.path.join(p, f'{rnd}_ncu_full_summary.txt'), 'w') as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout)
    for k in ('splat_depth', 'splat_feat', 'resolve'):
      f.write(f'\n-- warp stall samples, {k}\n')
      f.write(subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_stalls.py'), rep, k, '8'], capture_output=True, text=True).stdout)
  hdr, units, rows = raw_rows(rep)
  traffic = {}
  for r in rows:
    name = r[hdr.index('Kernel Name')].split('<')[0].split('(')[0].replace('void ', '').strip()
    rd, wr = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    traffic[name] = int(to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]))
  traffic['_note'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full` capture of '
                      'bench.py config c2 (cold caches: ncu flushes L2 between replays)')
  with open(os.path.join(p, 'traffic.json'), 'w') as f:
    json.dump(traffic, f, indent=1)
  # launch-list shares
  durs = {}
  for r in csv.reader(open(os.path.join(g, f'launches_{tag}.csv'))):
    if len(r) > 10 and 'gpu__time_duration.sum' in r:
      name = r[4].split('<')[0].replace('void ', '')
      val = next((x for x in reversed(r) if x.replace('.', '', 1).isdigit()), None)
      if val:
        durs.setdefault(name, []).append(float(val))
  tot = sum(sum(v) / len(v) for v in durs.values())
  with open(os.path.join(p, f'{rnd}_launch_shares.json'), 'w') as f:
    json.dump({k: {'mean_ns': sum(v) / len(v), 'launches': len(v), 'share': sum(v) / len(v) / tot} for k, v in durs.items()}, f, indent=1)
  for d in ('room', 'rand'):
    src = os.path.join(g, f'bench_{tag}_{d}.json')
    if os.path.exists(src):
      shutil.copy(src, os.path.join(p, f'{rnd}_bench_c2_{d}.json'))
  print(json.dumps(traffic, indent=1))
  print(open(os.path.join(p, f'{rnd}_launch_shares.json')).read())


if __name__ == '__main__':
  main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'r01')
