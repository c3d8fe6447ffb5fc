"""ORACLE (test infrastructure, not product code) -- ctypes loader for ref_exact.c.

Builds oracle/_build/libse3ds_oracle.so on first use (gcc, see oracle/Makefile)
and exposes the canonical-arithmetic restatement as numpy functions.  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import ref_numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libse3ds_oracle.so')
_lib = None
F32 = np.float32


def build(force: bool = False) -> str:
  src = os.path.join(_HERE, 'ref_exact.c')
  if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
    subprocess.check_call(['make', '-s', '-B', '-C', _HERE])
  return _SO


def lib():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(build())
  return _lib


def _p(a, t=ctypes.c_void_p):
  return None if a is None else a.ctypes.data_as(t)


def tables(h: int, w: int):
  """(elev, head, sin_e, cos_e, sin_h, cos_h) float32."""
  outs = [np.empty(h, F32), np.empty(w, F32), np.empty(h, F32), np.empty(h, F32), np.empty(w, F32), np.empty(w, F32)]
  lib().se3ds_oracle_tables(ctypes.c_int(h), ctypes.c_int(w), *[_p(o) for o in outs])
  return outs


def atan2f(y, x):
  y = np.ascontiguousarray(y, F32); x = np.ascontiguousarray(x, F32)
  out = np.empty_like(y)
  lib().se3ds_oracle_atan2f_array(_p(y), _p(x), _p(out), ctypes.c_longlong(y.size))
  return out


def acosf(q):
  q = np.ascontiguousarray(q, F32)
  out = np.empty_like(q)
  lib().se3ds_oracle_acosf_array(_p(q), _p(out), ctypes.c_longlong(q.size))
  return out


def unproject(depth: np.ndarray, depth_scale: float):
  """depth (N,H,W) f32 -> xyz1 (N,4,HW) f32, valid (N,HW) bool  (pano_utils.py:220-236)."""
  depth = np.ascontiguousarray(depth, F32)
  n, h, w = depth.shape
  xyz1 = np.empty((n, 4, h * w), F32)
  valid = np.empty((n, h * w), np.uint8)
  lib().se3ds_oracle_unproject(_p(depth), ctypes.c_int(n), ctypes.c_int(h), ctypes.c_int(w),
                               ctypes.c_float(depth_scale), _p(xyz1), _p(valid))
  return xyz1, valid.astype(bool)


def equirectangular_to_pointcloud(feats, depth, void_class, depth_scale, size_mult=1.0,
                                  interpolation_method='nearest'):
  """Canonical-arithmetic twin of ref_numpy.equirectangular_to_pointcloud."""
  # argument validation and feature handling are shared with the literal restatement
  _, ff = ref_numpy.equirectangular_to_pointcloud(feats, depth, void_class, depth_scale, size_mult,
                                                  interpolation_method)
  depth = np.asarray(depth, F32)
  if size_mult != 1.0:
    n, h, w = depth.shape
    depth = ref_numpy.tf_resize(depth[..., None], (int(h * size_mult), int(w * size_mult)), 'nearest')[..., 0]
  xyz1, _ = unproject(depth, depth_scale)
  return xyz1, ff


def splat(coords, feats, height, width, depth_scale, void_in, void_out=0.0, mode=0):
  """Canonical splat.  coords (N,4,M) f32; feats (N,M) or (N,M,C).
  Returns dict(depth, feat, zbuf, winner, flat, kept, rad)."""
  coords = np.ascontiguousarray(coords, F32)
  feats = np.asarray(feats)
  scalar = feats.ndim == 2
  if scalar:
    feats = feats[..., None]
  feats = np.ascontiguousarray(feats, F32)
  n, _, m = coords.shape
  c = feats.shape[-1]
  out = dict(depth=np.empty((n, height, width), F32), feat=np.empty((n, height, width, c), F32),
             zbuf=np.empty((n, height, width), F32), winner=np.empty((n, height, width), np.int32),
             flat=np.empty((n, m), np.int32), kept=np.empty((n, m), np.int32), rad=np.empty((n, m), F32),
             valid=np.empty((n, m), np.uint8))
  lib().se3ds_oracle_splat(_p(coords), _p(feats), ctypes.c_int(n), ctypes.c_longlong(m), ctypes.c_int(c),
                           ctypes.c_int(height), ctypes.c_int(width), ctypes.c_int(mode),
                           ctypes.c_float(void_in), ctypes.c_float(void_out), ctypes.c_float(depth_scale),
                           _p(out['depth']), _p(out['feat']), _p(out['zbuf']), _p(out['winner']),
                           _p(out['flat']), _p(out['kept']), _p(out['rad']), _p(out['valid']))
  if scalar:
    out['feat'] = out['feat'][..., 0]
  return out


def project_feats_to_equirectangular(feats, xyz1, height, width, void_class, depth_scale):
  o = splat(xyz1, feats, height, width, depth_scale, void_class, 0.0, mode=0)
  return o['depth'], o['feat']


def project_to_feat(transformed_coords, feats, height, width, depth_scale, input_void_class,
                    output_void_class=0):
  o = splat(transformed_coords, feats, height, width, depth_scale, input_void_class, output_void_class, mode=1)
  return o['depth'], o['feat']


def rotate(coords, rot):
  """coords (J,4,M), rot (J,3,3) -> rotated coords, canonical fma order (SE(3) extension)."""
  coords = np.ascontiguousarray(coords, F32)
  rot = np.ascontiguousarray(rot, F32).reshape(coords.shape[0], 9)
  out = np.empty_like(coords)
  lib().se3ds_oracle_rotate(_p(coords), _p(rot), ctypes.c_int(coords.shape[0]), ctypes.c_longlong(coords.shape[2]), _p(out))
  return out


def reproject(rgb, depth, src_pos, tgt_pos, depth_scale=ref_numpy.DEPTH_SCALE,
              unproject_void=ref_numpy.INVALID_RGB_VALUE, project_void=ref_numpy.INVALID_RGB_VALUE,
              mask_first_frame=True, mask_proportion=0.125, per_job_bin=False, tgt_rot=None):
  """Canonical twin of ref_numpy.reproject_trajectory, generalised to P target poses.

  rgb (N,S,H,W,3) int; depth (N,S,H,W); src_pos (N,S,3); tgt_pos (N,P,3) or (N,3).
  Jobs are ordered (n, p) row-major.  With per_job_bin=False the whole call is ONE
  reference call with batch N*P (global reject bin on job 0, pixel 0); with True
  every job is its own reference call (batch 1).
  Returns dict(image (J,H,W,3), depth (J,H,W,1), mask (J,H,W,1), winner (J,H,W), raw_rgb, zbuf, flat, kept).
  """
  rgb = np.asarray(rgb)
  n, s, h, w, _ = rgb.shape
  tgt_pos = np.asarray(tgt_pos, F32)
  if tgt_pos.ndim == 2:
    tgt_pos = tgt_pos[:, None, :]
  p = tgt_pos.shape[1]
  coords, feats = [], []
  for k in range(s):
    frame = rgb[:, k].astype(np.int32)
    if mask_first_frame and k == 0:
      frame = ref_numpy.mask_pano(frame, proportion=mask_proportion,
                                  masked_region_value=ref_numpy.INVALID_RGB_VALUE)
    xyz1, f = equirectangular_to_pointcloud(frame, depth[:, k], unproject_void, depth_scale)
    ps = np.concatenate([np.asarray(src_pos[:, k], F32), np.zeros((n, 1), F32)], axis=1)
    coords.append(xyz1 + ps[:, :, None])
    feats.append(f)
  coords = np.concatenate(coords, axis=2)          # (N,4,S*HW)
  feats = np.concatenate(feats, axis=1).astype(F32)  # (N,S*HW,3)
  t = np.concatenate([tgt_pos, np.zeros((n, p, 1), F32)], axis=2)  # (N,P,4)
  rel = (coords[:, None] - t[..., None]).astype(F32).reshape(n * p, 4, -1)
  if tgt_rot is not None:
    rel = rotate(rel, np.asarray(tgt_rot, F32).reshape(n * p, 3, 3))
  featj = np.broadcast_to(feats[:, None], (n, p) + feats.shape[1:]).reshape(n * p, -1, 3)
  if per_job_bin:
    outs = [splat(rel[j:j + 1], featj[j:j + 1], h, w, depth_scale, project_void) for j in range(n * p)]
    o = {k: np.concatenate([x[k] for x in outs], axis=0) for k in outs[0]}
  else:
    o = splat(rel, featj, h, w, depth_scale, project_void)
  image, d, mask = ref_numpy.guidance_from_projection(o['depth'], o['feat'])
  return dict(image=image, depth=d, mask=mask, winner=o['winner'], raw_rgb=o['feat'], zbuf=o['zbuf'],
              flat=o['flat'], kept=o['kept'], rad=o['rad'], valid=o['valid'], feats_in=featj)


def get_proportion_invalid_for_depth(position_offset, depth_image, distance_padding: float = 0.10):
  """Canonical twin of ref_numpy.get_proportion_invalid_for_depth (canonical atan2, IEEE sqrt)."""
  import math
  po = np.asarray(position_offset, F32)
  depth_image = np.asarray(depth_image, F32)
  distance = np.sqrt(F32(F32(po[0] * po[0]) + F32(po[1] * po[1])) + F32(po[2] * po[2]))
  height, width = depth_image.shape
  heading = atan2f(np.array([-po[0]], F32), np.array([-po[1]], F32))[0]
  if heading < 0:
    heading = F32(heading + F32(2 * math.pi))
  heading_proportion = F32(heading / F32(2 * math.pi))
  delta_xy = F32(math.sqrt(float(F32(po[1] * po[1]) + F32(po[0] * po[0]))))
  elevation = atan2f(np.array([delta_xy], F32), np.array([-po[2]], F32))[0]
  if elevation < 0:
    elevation = F32(elevation + F32(math.pi))
  elevation_proportion = F32(elevation / F32(math.pi))
  heading_start = int(F32(heading_proportion * F32(width)))
  elevation_start = int(F32(elevation_proportion * F32(height)))
  tw = int(30 / 360 * width)
  th = int(60 / 180 * height)
  region = depth_image[max(0, elevation_start - th):min(height, elevation_start + th),
                       max(0, heading_start - tw):min(width, heading_start + tw)]
  return float(F32(np.mean(region * F32(ref_numpy.DEPTH_SCALE) < F32(distance + F32(distance_padding)))))
