"""The kernels replace IEEE divisions by cheaper fma sequences (canon_math.cuh: div_const, div_rcp,
clip01_div255).  This test proves, on the CPU and with the same formulas in C, that the sequences
return exactly the IEEE quotient: exhaustively over all 2^23 mantissas for the constants
(2*pi, pi, 255, 20), and on random operands for the run-time denominator rad."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_division_sequences_are_exact():
  subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), '_build/divcheck'])
  out = subprocess.run([os.path.join(ROOT, 'oracle', '_build', 'divcheck'), '3', '30'], capture_output=True, text=True)
  assert out.returncode == 0, out.stdout
  assert out.stdout.count('div3_bad 0') == 6 and 'div5_bad 0' in out.stdout, out.stdout
  # the reciprocal constants compiled into the kernels are the correctly rounded ones
  text = open(os.path.join(ROOT, 'se3ds_b200', 'csrc', 'canon_math.cuh')).read()
  kern = open(os.path.join(ROOT, 'se3ds_b200', 'csrc', 'kernels.cuh')).read()
  ys = dict(re.findall(r'const ([0-9.]+) y=(\S+)', out.stdout))
  assert ys['6.28318548'] + 'f' in text and ys['3.14159274'] + 'f' in text and ys['255'] + 'f' in kern


def test_polynomial_coefficients_agree_between_oracle_and_kernels():
  """oracle/ref_exact.c and canon_math.cuh are written independently; their constants must match."""
  c = open(os.path.join(ROOT, 'oracle', 'ref_exact.c')).read()
  cu = open(os.path.join(ROOT, 'se3ds_b200', 'csrc', 'canon_math.cuh')).read()
  hexf = re.compile(r'-?0x1\.[0-9a-f]+p[+-]\d+f')
  atan_c = hexf.findall(c[c.index('kATAN[9]'):c.index('kASIN[6]')])
  asin_c = hexf.findall(c[c.index('kASIN[6]'):c.index('se3ds_oracle_atan2f')])
  assert len(atan_c) == 9 and len(asin_c) == 6
  atan_cu = hexf.findall(cu[cu.index('canon_atan_poly'):cu.index('canon_atan2f')])
  asin_cu = hexf.findall(cu[cu.index('canon_acosf'):cu.index('div_no_nan')])
  assert atan_cu[:9] == atan_c[::-1]
  assert asin_cu[:6] == asin_c[::-1]
