"""Float32 emulation (numpy) of the certified COLUMN of K2's fast projection
(se3ds_b200/csrc/canon_math.cuh project_pixel_fast, heading part; coefficients scaled as
se3ds_geom.cu does) and of the canonical column chain (ref_exact.c pseudo_perspective + pixel_of),
used by tests/test_column_certification.py to check the certificate adversarially:

  a certified column -- |fx_fast - rint(fx_fast)| > dx = W * 1e-6 -- must equal the canonical column,
  because |fx_fast - fx_canonical| < dx (worst-case bound derived in DESIGN.md section 4).

Points are PLANTED at column coordinates k +- dx (1 +- eps), over all eight octants, radii from 1e-3 to
1e4, and widths 10 ... 8192 (powers of two and not); the approximate reciprocal of the kernel (MUFU.RCP,
<= 1 ulp by the PTX ISA) is emulated as the correctly rounded reciprocal perturbed by -2 ... +2 ulp.

Not product code; nothing here is imported by the package.  Run: python tests/tools/col_cert_proto.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from oracle import ref_exact as E

F32 = np.float32
ATAN_C = [float.fromhex(x) for x in ('-0x1.dcc7b0p-10', '0x1.695cf0p-7', '-0x1.0126a6p-5', '0x1.dc8cccp-5', '-0x1.58b92ep-4',
                                     '0x1.c0c7a4p-4', '-0x1.242616p-3', '0x1.999266p-3', '-0x1.555540p-2')]
TWO_PI = F32(2 * np.pi)
PI15 = F32(1.5 * np.pi)
MARGIN = 1.0e-6


def fma32(a, b, c):
  """float32 fma emulated through float64 (the product of two float32 is exact in float64; the double
  rounding of the sum is far below the effects measured here)."""
  return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(F32)


def fast_fx(x, y, w, rcp_ulps=0):
  """project_pixel_fast, heading part -> fx (float32)."""
  kx = w / (2.0 * np.pi)
  ca = [F32(kx * c) for c in ATAN_C]
  ax, ay = np.abs(x), np.abs(y)
  mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
  with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
    rcp = ((F32(1) / mx).astype(F32) * F32(1 + rcp_ulps * 2.0 ** -23)).astype(F32)
    t = (mn * rcp).astype(F32)
    s = (t * t).astype(F32)
    p = np.full_like(s, ca[0])
    for c in ca[1:]:
      p = fma32(p, s, c)
    phi = (t * fma32(p, s, F32(kx))).astype(F32)
    phi = np.where(ay > ax, (F32(0.25 * w) - phi).astype(F32), phi)
    neg = np.signbit(y) ^ ~np.signbit(x)
    sphi = np.where(neg, -phi, phi)
    return (np.where(np.signbit(x), F32(0.25 * w), F32(0.75 * w)) + sphi).astype(F32)


def certified(fx, w, margin=MARGIN):
  with np.errstate(invalid='ignore'):
    return np.abs((fx - np.rint(fx)).astype(F32)) > F32(F32(w) * F32(margin))


def canonical_fx(x, y, z, w):
  """The canonical chain in float32 (every operation correctly rounded): rad, atan2, h, u, px, vx, fx."""
  rad = np.sqrt((((x * x).astype(F32) + (y * y).astype(F32)).astype(F32) + (z * z).astype(F32)).astype(F32)).astype(F32)
  h = (PI15 - E.atan2f(y, x)).astype(F32)
  h = np.where(h <= 0, (h + TWO_PI).astype(F32), h)
  h = np.where(h > TWO_PI, (h - TWO_PI).astype(F32), h)
  u = (((h / TWO_PI).astype(F32) * F32(2)).astype(F32) - F32(1)).astype(F32)
  px = (rad * u).astype(F32)
  vx = (px / rad).astype(F32)
  return ((((vx + F32(1)).astype(F32)) * F32(0.5)).astype(F32) * F32(w)).astype(F32)


def true_fx(x, y, w):
  """Exact column coordinate of the float32 point, in float64."""
  th = np.arctan2(y.astype(np.float64), x.astype(np.float64))
  return np.mod(w * (0.75 - th / (2 * np.pi)), w)


def planted_points(w, n_cols, eps_list, radii, rng):
  """Points whose exact column coordinate is k +- dx (1 + eps) for random columns k, for each radius."""
  dx = w * MARGIN
  ks = np.unique(np.concatenate([rng.integers(0, w, n_cols), [0, 1, w // 4, w // 2, 3 * w // 4, w - 1]]))
  fx = []
  for eps in eps_list:
    for sgn in (-1.0, 1.0):
      fx.append(ks + sgn * dx * (1.0 + eps))
  fx = np.mod(np.concatenate(fx), w)
  th = 2 * np.pi * (0.75 - fx / w)
  xs, ys, zs = [], [], []
  for r in radii:
    zfrac = rng.uniform(-0.9, 0.9, th.size)
    rho = r * np.sqrt(1 - zfrac ** 2)
    xs.append(rho * np.cos(th)); ys.append(rho * np.sin(th)); zs.append(r * zfrac)
  return (np.concatenate(xs).astype(F32), np.concatenate(ys).astype(F32), np.concatenate(zs).astype(F32))


def check(w, n_cols=3000, seed=0):
  """Returns (points, certified fraction, wrong certified, max |fast - canonical| / dx, max |canonical - exact| / dx)."""
  rng = np.random.default_rng(seed + w)
  eps_list = (-0.5, -0.1, -0.01, 0.01, 0.1, 0.5, 3.0)
  x, y, z = planted_points(w, n_cols, eps_list, (1e-3, 0.05, 1.0, 17.0, 300.0, 1e4), rng)
  # plus uniformly random directions and the |x| >> |y|, |y| >> |x| corners
  m = 200000
  th = rng.uniform(-np.pi, np.pi, m)
  r = 10.0 ** rng.uniform(-3, 4, m)
  x = np.concatenate([x, (r * np.cos(th)).astype(F32), (r * rng.uniform(-1e-4, 1e-4, m)).astype(F32)])
  y = np.concatenate([y, (r * np.sin(th)).astype(F32), (r * np.sign(rng.uniform(-1, 1, m))).astype(F32)])
  z = np.concatenate([z, (r * rng.uniform(-1, 1, m)).astype(F32), (r * rng.uniform(-1, 1, m)).astype(F32)])
  fc = canonical_fx(x, y, z, w)
  col_c = np.floor(fc).astype(np.int64)
  ok = np.isfinite(fc)
  dx = w * MARGIN
  wrong = 0
  dev = 0.0
  frac = 1.0
  for ulps in (-2, -1, 0, 1, 2):
    ff = fast_fx(x, y, w, ulps)
    cert = certified(ff, w) & ok
    wrong += int(np.sum(np.floor(ff[cert]).astype(np.int64) != col_c[cert]))
    # the wrap-around at fx = W / 0 is a difference of W between neighbours, not an error
    d = np.abs(ff.astype(np.float64) - fc.astype(np.float64))
    d = np.minimum(d, np.abs(d - w))
    dev = max(dev, float(np.nanmax(d[ok])))
    frac = min(frac, float(cert.mean()))
  d0 = np.abs(fc.astype(np.float64) - true_fx(x, y, w))
  d0 = np.minimum(d0, np.abs(d0 - w))
  return x.size, frac, wrong, dev / dx, float(np.nanmax(d0[ok])) / dx


if __name__ == '__main__':
  print('    W     points  certified  wrong   max|fast-canon|/dx  max|canon-exact|/dx')
  for w in (10, 126, 1000, 1024, 2048, 4096, 8192):
    n, frac, wrong, dev, dev0 = check(w)
    print(f'{w:6d} {n:9d}   {frac * 100:6.2f} %  {wrong:5d}      {dev:8.3f}            {dev0:8.3f}')
