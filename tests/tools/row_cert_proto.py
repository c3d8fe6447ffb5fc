"""CPU prototype (numpy float32) of a cheaper certified ROW computation for K2's fast projection.

Idea (profiles/r01_summary.md, "Not tried yet"): instead of the full asin polynomial + reduction,
estimate e = acos(q), q = z / rad, with a short sqrt(1-|q|) * poly(|q|) approximation, take
row = floor(e * H / pi), and CERTIFY the row by comparing q with the tabulated cosines of the two row
boundaries (cos is monotone) with a constant margin.  This script measures, against the bit-exact
oracle, (a) that no certified row differs from the canonical row, (b) the deferral rate.

Not product code; nothing here is imported by the package.  Run: python tests/tools/row_cert_proto.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from oracle import ref_exact as E, ref_numpy as R
from se3ds_b200 import synth

F32 = np.float32


def fit_acos_poly(ncoef):
  """acos(x) ~ sqrt(1 - x) * (c0 + c1 x + ... ) on [0, 1] (Abramowitz-Stegun form), least squares on
  Chebyshev nodes in float64, coefficients rounded to float32."""
  k = np.arange(4000)
  x = 0.5 - 0.5 * np.cos(np.pi * (k + 0.5) / 4000)
  x = x[x < 0.9999]
  target = np.arccos(x) / np.sqrt(1 - x)
  A = np.vander(x, ncoef, increasing=True)
  c, *_ = np.linalg.lstsq(A, target, rcond=None)
  return c.astype(F32)


def approx_acos(q, c):
  a = np.abs(q).astype(F32)
  p = np.full_like(a, c[-1])
  for ci in c[-2::-1]:
    p = (p * a + ci).astype(F32)
  r = (np.sqrt((F32(1) - a).astype(F32)).astype(F32) * p).astype(F32)
  return np.where(q < 0, (F32(np.pi) - r).astype(F32), r).astype(F32)


def run(dist, h, ncoef, mq, seed=0, rcp_ulps=2):
  w = 2 * h
  inp = synth.make_inputs(1, 1, 1, h, seed=seed, dist=dist)
  rgb = R.mask_pano(inp['rgb'][0, 0].astype(np.int32)[None], 0.125, -1)
  xyz1, feats = E.equirectangular_to_pointcloud(rgb, inp['depth'][0:1, 0], -1, 20.0)
  xyz = ((xyz1[0, :3] + inp['src_pos'][0, 0][:, None]).astype(F32) - inp['tgt_pos'][0, 0][:, None]).astype(F32)
  out = E.splat(np.concatenate([xyz, np.ones((1, xyz.shape[1]), F32)])[None], feats, h, w, 20.0, -1.0, 0.0, 0)
  valid = out['valid'][0].astype(bool)
  row_canon = out['flat'][0] // w
  x, y, z = xyz
  rad = out['rad'][0]
  c = fit_acos_poly(ncoef)
  cosb = np.cos(np.arange(h + 1) * np.pi / h).astype(F32)          # row boundaries, decreasing
  worst = 0
  stats = []
  for sgn in (-1, 0, 1):                                            # approximate reciprocal: +-rcp_ulps ulp
    rinv = (F32(1) / rad).astype(F32)
    rinv = (rinv * F32(1 + sgn * rcp_ulps * 2.0 ** -23)).astype(F32)
    q = (z * rinv).astype(F32)
    e = approx_acos(q, c)
    fy = (e * F32(h / np.pi)).astype(F32)
    row = np.floor(fy).astype(np.int64)
    inside = (row >= 0) & (row < h)
    rc = np.clip(row, 0, h - 1)
    certain = inside & (q < cosb[rc] - F32(mq)) & (q > cosb[rc + 1] + F32(mq)) & np.isfinite(q)
    sel = valid & certain
    bad = int(np.sum(row[sel] != row_canon[sel]))
    worst = max(worst, bad)
    stats.append(1.0 - certain[valid].mean())
  return worst, max(stats), float(np.abs(approx_acos(np.linspace(-1, 1, 200001).astype(F32), c) - np.arccos(np.linspace(-1, 1, 200001))).max())


if __name__ == '__main__' and len(sys.argv) == 1:
  print('dist    H   coef  margin_q   wrong-certified  deferred   max |acos err| (rad)')
  for dist in ('room', 'rand'):
    for h in (512, 2048) if dist == 'room' else (512,):
      for ncoef in (4, 5, 6):
        for mq in (6.3e-6, 3e-6):
          bad, deferred, err = run(dist, h, ncoef, mq)
          print(f'{dist:5s} {h:5d}  {ncoef:4d}  {mq:8.1e}   {bad:8d}        {deferred * 100:6.2f} %   {err:.2e}')


# ------------------------------------------------------------------------------------------
# Faithful float32 emulation of the kernel code (canon_math.cuh project_pixel_fast, row part) and of the
# host table (se3ds_geom.cu get_tables), with the approximate reciprocal / square root perturbed by
# +-2 ulp.  `python tests/tools/row_cert_proto.py sweep` checks many shapes and seeds.
KCOEF = np.array([float.fromhex(x) for x in ('0x1.921f16p+0', '-0x1.b67528p-3', '0x1.5a1b66p-4', '-0x1.22be94p-5', '0x1.171b8cp-7')], F32)


def host_table(h, margin_scale=1e-6):
  m = 2.0 * np.pi * margin_scale
  r = np.arange(h)
  lo = (np.cos((r + 1) * np.pi / h) + m).astype(F32)
  hi = (np.cos(r * np.pi / h) - m).astype(F32)
  return np.nextafter(lo, F32(2)), np.nextafter(hi, F32(-2))


def kernel_rows(z, rad, h, sr, ss):
  """canon_math.cuh project_pixel_fast, row part: q = z * rsqrt(r2) (the approximate reciprocal square root
  of the squared radius doubles as 1 / rad; emulated as 1 / rad off by sr * 2 ulp), candidate row from the
  polynomial in row units (coefficients times H / pi, as se3ds_geom.cu scales them), certified against the
  cosine table."""
  rinv = ((F32(1) / rad).astype(F32) * F32(1 + sr * 2.0 ** -22)).astype(F32)
  q = (z * rinv).astype(F32)
  a = np.abs(q)
  ce = (KCOEF.astype(np.float64) * (h / np.pi)).astype(F32)  # lowest degree first here
  p = np.full_like(a, ce[4])
  for c in ce[3::-1]:
    p = (p * a + c).astype(F32)          # fma in the kernel: one rounding less, irrelevant at 1e-5
  with np.errstate(invalid='ignore'):
    sq = (np.sqrt((F32(1) - a).astype(F32)) * F32(1 + ss * 2.0 ** -22)).astype(F32)
  fy = (sq * p).astype(F32)
  fy = np.where(q < 0, (F32(h) - fy).astype(F32), fy)
  with np.errstate(invalid='ignore'):
    row = np.where(np.isfinite(fy), np.floor(fy), 0).astype(np.int64)
  row = np.where(row < 0, h - 1, np.minimum(row, h - 1))  # min((unsigned)floor, H - 1)
  lo, hi = host_table(h)
  certain = (q > lo[row]) & (q < hi[row])
  return row, certain


def sweep():
  total = wrong = 0
  worst_deferred = {}
  for h in (3, 4, 5, 8, 16, 33, 64, 128, 256, 512, 1024, 2048):
    seeds = range(6) if h <= 512 else range(2)
    for seed in seeds:
      for dist in ('room', 'rand'):
        w = 2 * h
        inp = synth.make_inputs(1, 1, 1, h, seed=seed, dist=dist, sweep=bool(seed & 1))
        rgb = inp['rgb'][0, 0].astype(np.int32)[None]
        xyz1, feats = E.equirectangular_to_pointcloud(rgb, inp['depth'][0:1, 0], -1, 20.0)
        xyz = ((xyz1[0, :3] + inp['src_pos'][0, 0][:, None]).astype(F32) - inp['tgt_pos'][0, 0][:, None]).astype(F32)
        out = E.splat(np.concatenate([xyz, np.ones((1, xyz.shape[1]), F32)])[None], feats, h, w, 20.0, -1.0, 0.0, 0)
        valid = out['valid'][0].astype(bool)
        row_canon = out['flat'][0] // w
        for sr in (-1, 0, 1):
          for ss in (-1, 0, 1):
            row, certain = kernel_rows(xyz[2], out['rad'][0], h, sr, ss)
            sel = valid & certain
            wrong += int(np.sum(row[sel] != row_canon[sel]))
            total += int(sel.sum())
            # points the canonical path rejects (row out of range etc.) must never be certified into the image
            inval = (~valid) & certain & (out['rad'][0] > 0)
            # (~valid also holds void-feature points; only geometric rejections matter here: none exist for e in [0, pi])
            worst_deferred[h] = max(worst_deferred.get(h, 0.0), 1.0 - certain[valid].mean())
  print('certified points checked:', total, ' wrong rows:', wrong)
  print('worst deferred fraction per H:', {k: round(v * 100, 2) for k, v in worst_deferred.items()})


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'sweep':
  sweep()
