"""Writes profiles/<round>_sass_summary.txt: resource usage and SASS mnemonic counts (cuobjdump) of the kernels the
c2 bench line launches.  Runs wherever the CUDA toolkit is installed; no GPU needed.  Usage: sass_summary.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'se3ds_b200', 'lib', 'libse3ds_geom.so')
WANT = {
    '_ZN5se3ds18splat_depth_kernelIhLb1ELb1ELi1ELb0ELb0ELb1EEEvNS_11FusedParamsE': 'K2 splat_depth_kernel<uint8, VEC, FAST, PROJ=1, key32, no rotation, PLAIN>',
    '_ZN5se3ds18splat_depth_kernelIhLb1ELb1ELi1ELb1ELb0ELb1EEEvNS_11FusedParamsE': 'K2 splat_depth_kernel<..., key64, ..., PLAIN>',
    '_ZN5se3ds17splat_feat_kernelIhLi4ELb0EEEvNS_11FusedParamsE': 'K3 splat_feat_kernel<uint8, 4 points per thread, key32>',
    '_ZN5se3ds14resolve_kernelILi4ELb0ELb0EEEvNS_11FusedParamsE': 'K4 resolve_kernel<4 pixels per thread, key32, float32 outputs>',
    '_ZN5se3ds14resolve_kernelILi4ELb0ELb1EEEvNS_11FusedParamsE': 'K4 resolve_kernel<4 pixels per thread, key32, compact outputs>',
}
SPECIAL = ('REDG', 'ATOMG', 'LDG', 'STG', 'MUFU', 'VOTE', 'STS', 'LDS', 'HADD2', 'PRMT', 'ACQBULK', 'BAR', 'F2I', 'FRND')


def main(rnd):
  sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
  res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
  usage, name = {}, None
  for line in res.splitlines():
    m = re.match(r'\s*Function (\S+):', line)
    if m:
      name = m.group(1)
    elif name and 'REG:' in line:
      usage[name], name = line.strip(), None
  out = ['# cuobjdump of se3ds_b200/lib/libse3ds_geom.so (sm_100a): resource usage and SASS mnemonic counts of the kernels the',
         '# c2 bench line launches (static counts; per-launch dynamic counts are in the ncu summaries).', '']
  for blk in sass.split('\t\tFunction : ')[1:]:
    fn = blk.split('\n', 1)[0].strip()
    if fn not in WANT:
      continue
    ops = collections.Counter()
    for line in blk.split('\n'):
      m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
      if m:
        ops[m.group(1)] += 1
    out += ['== ' + WANT[fn], '   ' + fn, '   ' + usage.get(fn, '(usage not found)'),
            '   %d SASS instructions; local memory ops (LDL/STL): %d' % (sum(ops.values()), sum(v for k, v in ops.items() if k.startswith(('LDL', 'STL')))),
            '   memory / special: ' + ', '.join('%s x%d' % (k, ops[k]) for k in sorted(ops) if k.startswith(SPECIAL)),
            '   most frequent: ' + ', '.join('%s x%d' % kv for kv in ops.most_common(12)), '']
  with open(os.path.join(ROOT, 'profiles', rnd + '_sass_summary.txt'), 'w') as f:
    f.write('\n'.join(out))


if __name__ == '__main__':
  main(sys.argv[1] if len(sys.argv) > 1 else 'r02')
