"""Bounds the damage of the unpinned bit authority.

TensorFlow cannot run here, so the kernels are bit-compared with oracle/ref_exact.c, whose atan2 / acos /
sin / cos are canonical float32 definitions, not TF's libm.  Two float32 pipelines that differ by an ulp in
a transcendental can only disagree on a pixel index where the coordinate sits within a few error units of
a pixel border, one error unit being one ulp of the coordinate plus the first-order effect on it of one
ulp of the point's cartesian coordinates (the sin / cos tables of the unprojection differ by an ulp
between the pipelines; next to the target's vertical axis, where the heading is ill-conditioned, that
is amplified by |p| / rho).  This test proves that statement for the two pipelines we do have --
ref_exact (canonical) and ref_numpy (numpy's libm) -- on full-size panoramas: EVERY point whose target
pixel differs lies within `MAX_UNITS` error units of an integer coordinate (measured maximum: printed,
asserted), and the disagreement rate is reported.  Real TensorFlow, the day it is importable, is compared the same way in
tests/test_tf_reference.py."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import ref_exact as X
from oracle import ref_numpy as R
from se3ds_b200 import synth

F32 = np.float32
_spec = importlib.util.spec_from_file_location(
    'col_cert_proto', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools', 'col_cert_proto.py'))
proto = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proto)

MAX_UNITS = 16.0


def canonical_fy(x, y, z, h):
  rad = np.sqrt((((x * x).astype(F32) + (y * y).astype(F32)).astype(F32) + (z * z).astype(F32)).astype(F32)).astype(F32)
  with np.errstate(divide='ignore', invalid='ignore'):
    q = np.where(rad == 0, F32(0), (z / rad).astype(F32)).astype(F32)
    e = X.acosf(q)
    v = (((e / F32(np.pi)).astype(F32) * F32(2)).astype(F32) - F32(1)).astype(F32)
    py = (rad * v).astype(F32)
    vy = (py / rad).astype(F32)
    return ((((vy + F32(1)).astype(F32)) * F32(0.5)).astype(F32) * F32(h)).astype(F32)


def border_distance_units(f, cond):
  """Distance of a coordinate from the nearest integer in error units: ulp(f) + cond, cond = the pixel
  shift caused by a one-ulp perturbation of the point's cartesian coordinates."""
  f64 = f.astype(np.float64)
  return np.abs(f64 - np.rint(f64)) / (np.spacing(np.maximum(np.abs(f), F32(1))).astype(np.float64) + cond)


def disagreement(h, dist, seed):
  w = 2 * h
  inp = synth.make_inputs(1, 1, 1, h, seed=seed, dist=dist)
  rgb = inp['rgb'].astype(np.int32)
  ox = X.reproject(rgb, inp['depth'], inp['src_pos'], inp['tgt_pos'], mask_first_frame=False)
  _, _, _, dbg = R.reproject_trajectory(rgb, inp['depth'], inp['src_pos'], inp['tgt_pos'][:, 0], mask_first_frame=False)
  fx_, fn_ = ox['flat'].reshape(-1), dbg['flat'].reshape(-1)
  bad = np.nonzero(fx_ != fn_)[0]
  xyz1, _ = X.equirectangular_to_pointcloud(rgb[:, 0], inp['depth'][:, 0], -1, 20.0)
  xyz = ((xyz1[0, :3] + inp['src_pos'][0, 0][:, None]).astype(F32) - inp['tgt_pos'][0, 0][:, None]).astype(F32)
  x, y, z = xyz[:, bad]
  fx = proto.canonical_fx(x, y, z, w)
  fy = canonical_fy(x, y, z, h)
  # one ulp of the largest term that went into the coordinates (local point, source and target position)
  mag = np.maximum(np.abs(xyz1[0, :3, bad]).max(axis=1), max(np.abs(inp['src_pos']).max(), np.abs(inp['tgt_pos']).max()))
  eps = np.spacing(mag.astype(F32)).astype(np.float64)
  rho = np.sqrt(x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2)
  rad = np.sqrt(rho ** 2 + z.astype(np.float64) ** 2)
  with np.errstate(divide='ignore'):
    cond_x = (w / (2 * np.pi)) * eps / rho
    cond_y = (h / np.pi) * eps / rad
  d = np.minimum(border_distance_units(fx, cond_x), border_distance_units(fy, cond_y))
  return fx_.size, bad.size, (float(d.max()) if bad.size else 0.0)


@pytest.mark.parametrize('h,dist,seed', [(512, 'room', 0), (512, 'rand', 1), (2048, 'room', 2)])
def test_every_disagreeing_point_sits_on_a_pixel_border(h, dist, seed):
  total, bad, worst = disagreement(h, dist, seed)
  rate = bad / total
  print(f'\n{h}x{2 * h} {dist}: ref_exact vs ref_numpy pixel indices differ for {bad} of {total} points '
        f'({rate:.2e}); farthest such point is {worst:.1f} error units from a pixel border')
  assert rate <= 2e-4, rate
  assert worst <= MAX_UNITS, worst
