"""ORACLE (test infrastructure, not product code) -- literal numpy restatement.

This module restates, op for op and in float32, the reference's geometric
guidance path so that the CUDA product path can be checked against it.  It is
the *semantic authority*: every function mirrors one reference function and
keeps its operation order, its float32 roundings, its trunc-toward-zero casts,
its `divide_no_nan`, its global reject bin at flat index 0, its +0.1 m
tolerance and its per-channel scatter-max.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this module.  Nothing under `se3ds_b200/` may.

Parity status: TensorFlow cannot be imported in this image, so the TF outputs
themselves are not available ("bit-level parity unpinned").  What *is* pinned:
the reference's own known-answer tests, restated in tests/test_oracle_kat.py
  * models/models_test.py:81-137   (plane KAT, 24 memory columns, y==1 / x==2)
  * models/models_test.py:64-68    (identity re-projection >= 95 % RGB equal)
  * inference/perturbation_utils_test.py:30-94 (five KATs)
  * utils/pano_utils_test.py:35-65 (3x6 ray table)
TF op semantics restated here (TF 2.8 sources, not executed):
  tf.linspace  -> [start, start + delta*i (i=1..n-2), stop], delta=(stop-start)/(n-1)
  tf.cast(f32->int32) -> truncation toward zero, NaN / out of range -> INT_MIN
  tf.math.divide_no_nan(x, 0) -> 0
  tf.tensor_scatter_nd_min/max -> order independent reductions incl. the init
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional, Tuple

import numpy as np

F32 = np.float32

# constants.py:21-29
INVALID_SEM_VALUE = 0
INVALID_RGB_VALUE = -1
PI = 3.1415926535897932384626433
HFOV = 90 * PI / 180
DEPTH_SCALE = 20.0
NUM_MP3D_CLASSES = 42
PANO_VIDEO_LENGTH = 8


# --------------------------------------------------------------------------
# TF primitives
# --------------------------------------------------------------------------
def tf_linspace(start, stop, num: int, dtype=np.float32) -> np.ndarray:
  """tf.linspace (math_ops.linspace_nd): exact end points, start+delta*i inside."""
  start = dtype(start)
  stop = dtype(stop)
  if num == 1:
    return np.array([start], dtype=dtype)
  n_steps = max(num - 1, 1)
  delta = dtype((stop - start) / dtype(n_steps))
  inner = (start + delta * np.arange(1, num - 1).astype(dtype)).astype(dtype)
  return np.concatenate([[start], inner, [stop]]).astype(dtype)


def divide_no_nan(x: np.ndarray, y: np.ndarray) -> np.ndarray:
  """tf.math.divide_no_nan: 0 where the denominator is 0."""
  y = np.broadcast_to(y, x.shape)
  out = np.zeros_like(x)
  np.divide(x, y, out=out, where=(y != 0))
  return out


def cast_int32(x: np.ndarray) -> np.ndarray:
  """tf.cast(float32 -> int32): trunc toward zero; NaN / overflow -> INT_MIN (x86)."""
  with np.errstate(invalid='ignore'):
    bad = ~np.isfinite(x) | (x >= 2147483648.0) | (x < -2147483648.0)
    out = np.where(bad, 0, x).astype(np.int32)
  out[bad] = np.iinfo(np.int32).min
  return out


def scatter_min(init: np.ndarray, idx: np.ndarray, upd: np.ndarray) -> np.ndarray:
  """tf.tensor_scatter_nd_min on a flat (P, C) buffer with (K,) indices."""
  out = init.copy()
  np.minimum.at(out, idx, upd)
  return out


def scatter_max(init: np.ndarray, idx: np.ndarray, upd: np.ndarray) -> np.ndarray:
  """tf.tensor_scatter_nd_max on a flat (P, C) buffer with (K,) indices."""
  out = init.copy()
  np.maximum.at(out, idx, upd)
  return out


# --------------------------------------------------------------------------
# utils/pano_utils.py
# --------------------------------------------------------------------------
def mask_pano(pano: np.ndarray, proportion: float = 0.125, masked_region_value=0) -> np.ndarray:
  """utils/pano_utils.py:245-265.  Rows [mh, H-mh] are kept (note the <=)."""
  _, height, _, _ = pano.shape
  masked_height = int(height * proportion)
  rng = np.arange(0, height)
  mask = np.logical_and(rng >= masked_height, rng <= height - masked_height)
  mask = mask.astype(pano.dtype)[None, :, None, None]
  one = np.array(1, dtype=pano.dtype)
  mrv = np.array(masked_region_value).astype(pano.dtype)
  return (mask * pano + (one - mask) * mrv).astype(pano.dtype)


def equirect_angle_tables(height: int, width: int) -> Tuple[np.ndarray, np.ndarray]:
  """elevation (H,), heading (W,) of utils/pano_utils.py:211-219 (float32)."""
  half_pixel_width = 0.5 * np.pi / height
  elevation = tf_linspace(half_pixel_width, np.pi - half_pixel_width, height)
  heading = tf_linspace(1.5 * np.pi - half_pixel_width, -0.5 * np.pi + half_pixel_width, width)
  return elevation, heading


def tf_resize(images: np.ndarray, size, method: str) -> np.ndarray:
  """tf.image.resize (TF 2.x: half_pixel_centers=True) for (N,H,W,C).
  'nearest': in = min(floor((out + 0.5) * scale), in_size - 1), dtype kept
  (resize_nearest_neighbor_op.cc, HalfPixelScalerForNN).  'bilinear': in = (out + 0.5) * scale - 0.5,
  lower = max(floor(in), 0), upper = min(ceil(in), in_size - 1), lerp = in - floor(in); float32 output
  (resize_bilinear_op.cc, compute_interpolation_weights).  scale = in_size / out_size in float32."""
  images = np.asarray(images)
  n, h, w, c = images.shape
  oh, ow = int(size[0]), int(size[1])
  sy, sx = F32(h) / F32(oh), F32(w) / F32(ow)
  if method == 'nearest':
    iy = np.minimum(np.floor((np.arange(oh, dtype=F32) + F32(0.5)) * sy), h - 1).astype(np.int64)
    ix = np.minimum(np.floor((np.arange(ow, dtype=F32) + F32(0.5)) * sx), w - 1).astype(np.int64)
    return images[:, iy][:, :, ix]
  if method != 'bilinear':
    raise NotImplementedError(method)
  def weights(out_size, in_size, scale):
    pos = ((np.arange(out_size, dtype=F32) + F32(0.5)) * scale - F32(0.5)).astype(F32)
    fl = np.floor(pos)
    lower = np.maximum(fl, 0).astype(np.int64)
    upper = np.minimum(np.ceil(pos), in_size - 1).astype(np.int64)
    return lower, upper, (pos - fl).astype(F32)
  ylo, yhi, yl = weights(oh, h, sy)
  xlo, xhi, xl = weights(ow, w, sx)
  img = images.astype(F32)
  tl = img[:, ylo][:, :, xlo]; tr = img[:, ylo][:, :, xhi]
  bl = img[:, yhi][:, :, xlo]; br = img[:, yhi][:, :, xhi]
  xl = xl[None, None, :, None]; yl = yl[None, :, None, None]
  top = tl + (tr - tl) * xl
  bottom = bl + (br - bl) * xl
  return (top + (bottom - top) * yl).astype(F32)


def tf_resize_antialias_triangle(images: np.ndarray, size) -> np.ndarray:
  """tf.image.resize(..., 'bilinear', antialias=True) == ScaleAndTranslate with the triangle kernel
  (TF 2.8 core/kernels/image/scale_and_translate_op.cc, ComputeSpansCore + GatherSpans; not vendored
  in /root/reference, restated from the published algorithm -- parity with TF's bits unpinned).
  Per output coordinate x: sample = (x + 0.5) * in/out, kernel_scale = max(in/out, 1), the span covers
  the source pixels whose centre lies within kernel_scale of the sample, clipped to the image, and the
  triangle weights are normalised by their sum.  Used by crop_pano(resize_to_original=True),
  utils/pano_utils.py:299-301.  Rows first, then columns, float32 throughout."""
  images = np.asarray(images)
  n, h, w, c = images.shape
  oh, ow = int(size[0]), int(size[1])
  def spans(in_size, out_size):
    inv_scale = F32(in_size) / F32(out_size)
    ks = max(inv_scale, F32(1.0))
    mat = np.zeros((out_size, in_size), dtype=F32)
    for x in range(out_size):
      sample = (F32(x) + F32(0.5)) * inv_scale
      lo = int(np.ceil(sample - ks - F32(0.5)))
      hi = int(np.floor(sample + ks - F32(0.5)))
      lo, hi = max(lo, 0), min(hi, in_size - 1)
      src = np.arange(lo, hi + 1, dtype=F32)
      wts = np.maximum(F32(0), F32(1) - np.abs((src + F32(0.5) - sample) / ks)).astype(F32)
      tot = F32(wts.sum(dtype=F32))
      if abs(tot) >= 1000 * np.finfo(F32).tiny:
        wts = (wts * (F32(1) / tot)).astype(F32)
      mat[x, lo:hi + 1] = wts
    return mat
  my, mx = spans(h, oh), spans(w, ow)
  img = images.astype(F32)
  rows = np.einsum('oh,nhwc->nowc', my, img).astype(F32)
  return np.einsum('pw,nowc->nopc', mx, rows).astype(F32)


def crop_pano(pano: np.ndarray, proportion: float = 0.125, method: str = 'bilinear',
              resize_to_original: bool = False) -> np.ndarray:
  """utils/pano_utils.py:268-303: crop int(H*p) rows top and bottom, optionally resize back
  (antialiased; 'nearest' ignores antialias), cast to the input dtype (truncation for integers)."""
  pano = np.asarray(pano)
  squeeze = pano.ndim == 3
  x = pano[None] if squeeze else pano
  if x.ndim != 4:
    raise ValueError(f'pano should be of shape (N, H, W, C), got {pano.shape} instead.')
  _, h, w, _ = x.shape
  mh = int(h * proportion)
  x = x[:, mh:h - mh]
  if resize_to_original:
    x = tf_resize(x, (h, w), 'nearest') if method == 'nearest' else tf_resize_antialias_triangle(x, (h, w))
  x = x.astype(pano.dtype)
  return x[0] if squeeze else x


def equirectangular_to_pointcloud(feats: np.ndarray, depth: np.ndarray, void_class,
                                  depth_scale: float, size_mult: float = 1.0,
                                  interpolation_method: str = 'nearest'):
  """utils/pano_utils.py:164-242 (size_mult == 1.0 only: resize is the identity).

  'bilinear' turns the features into float32 (tf.image.resize returns f32),
  'nearest' keeps their dtype.
  """
  if feats.ndim != 3 and feats.ndim != 4:
    raise ValueError('feats should have shape (N, H, W) or (N, H, W, C),'
                     f' got {feats.shape} instead.')
  if void_class < 0.0 and feats.dtype in (np.uint8, np.uint16, np.uint32, np.uint64):
    raise ValueError('feats datatype must be signed if the void class is negative')
  is_scalar_feat = feats.ndim == 3
  if is_scalar_feat:
    feats = feats[..., None]
  batch_size, height, width, channels = feats.shape
  assert width == 2 * height, 'Expected equirectangular input images'
  pano_depth = depth.astype(F32)
  pano_feats = feats.astype(F32) if interpolation_method != 'nearest' else feats
  if size_mult != 1.0:  # pano_utils.py:203-208: depth is resized 'nearest', features with the given method
    height, width = int(height * size_mult), int(width * size_mult)
    pano_depth = tf_resize(pano_depth[..., None], (height, width), 'nearest')[..., 0]
    pano_feats = tf_resize(feats, (height, width), interpolation_method)
  elevation, heading = equirect_angle_tables(height, width)
  depth_mask = np.logical_and(pano_depth > 0, pano_depth < F32(1.0)).astype(F32)
  rad = (pano_depth * F32(depth_scale)) * depth_mask
  void = np.array(void_class).astype(pano_feats.dtype)
  pano_feats = np.where(depth_mask[..., None] == 0, void, pano_feats)
  sin_e = np.sin(elevation)[None, :, None]
  cos_e = np.cos(elevation)[None, :, None]
  sin_h = np.sin(heading)[None, None, :]
  cos_h = np.cos(heading)[None, None, :]
  x = rad * sin_e * cos_h
  y = rad * sin_e * sin_h
  z = rad * cos_e
  xyz1 = np.stack([
      x.reshape(batch_size, -1), y.reshape(batch_size, -1), z.reshape(batch_size, -1),
      np.ones((batch_size, height * width), dtype=F32)], axis=1).astype(F32)
  filtered_feats = pano_feats.reshape(batch_size, -1, channels)
  if is_scalar_feat:
    filtered_feats = filtered_feats[..., 0]
  return xyz1, filtered_feats


def equirect_pseudo_perspective(xyz1: np.ndarray) -> np.ndarray:
  """utils/pano_utils.py:139-156: cartesian -> (rad*u, rad*v, rad, 1)."""
  xyz1 = xyz1.astype(F32)
  x, y, z = xyz1[:, 0, :], xyz1[:, 1, :], xyz1[:, 2, :]
  rad = np.sqrt(x * x + y * y + z * z)  # tf.pow(., 0.5)
  heading = np.arctan2(y, x)
  heading = F32(1.5 * math.pi) - heading
  heading = heading + F32(2 * math.pi) * (heading <= 0).astype(F32)
  heading = heading - F32(2 * math.pi) * (heading > F32(2 * math.pi)).astype(F32)
  with np.errstate(invalid='ignore'):
    elevation = np.arccos(divide_no_nan(z, rad))
  proj_x = rad * ((heading / F32(2 * math.pi)) * F32(2) - F32(1))
  proj_y = rad * ((elevation / F32(math.pi)) * F32(2) - F32(1))
  return np.stack([proj_x, proj_y, rad, np.ones_like(proj_x)], axis=1).astype(F32)


def project_feats_to_equirectangular(feats, xyz1, height, width, void_class, depth_scale,
                                     return_debug: bool = False):
  """utils/pano_utils.py:117-161."""
  proj_xyz1 = equirect_pseudo_perspective(xyz1)
  return project_to_feat(proj_xyz1, np.asarray(feats).astype(F32), height, width,
                         depth_scale=depth_scale, input_void_class=void_class,
                         return_debug=return_debug)


# --------------------------------------------------------------------------
# utils/point_cloud_utils.py
# --------------------------------------------------------------------------
def get_intrinsic_matrix(hfov: float) -> np.ndarray:
  """utils/point_cloud_utils.py:23-29."""
  return np.array([
      [1 / np.tan(hfov / 2.), 0., 0., 0.],
      [0., 1 / np.tan(hfov / 2.), 0., 0.],
      [0., 0., 1, 0],
      [0., 0., 0, 1]], dtype=F32)


def get_filtered_coords_and_feats(feats, depth, depth_scale):
  """utils/point_cloud_utils.py:32-87 (legacy perspective unprojection)."""
  if feats.ndim != 3 and feats.ndim != 4:
    raise ValueError('feats should have shape (N, H, W) or (N, H, W, C),'
                     f' got {feats.shape} instead.')
  is_scalar_feat = feats.ndim == 3
  if is_scalar_feat:
    feats = feats[..., None]
  batch_size, height, width = depth.shape
  channels = feats.shape[-1]
  # tf.linspace(-1, 1, n) with Python ints is evaluated in float64, then cast.
  xs_1d = tf_linspace(-1, 1, width, dtype=np.float64).astype(F32)
  ys_1d = tf_linspace(-1, 1, height, dtype=np.float64).astype(F32)
  xs, ys = np.meshgrid(xs_1d, ys_1d)
  xs = np.broadcast_to(xs.reshape(1, 1, height, width), (batch_size, 1, height, width))
  ys = np.broadcast_to(ys.reshape(1, 1, height, width), (batch_size, 1, height, width))
  depth = (depth.astype(F32) * F32(depth_scale))[:, None, :, :]
  ones = np.ones_like(depth)
  xyz = np.concatenate([xs * depth, ys * depth, depth, ones], axis=1)
  depth = depth.reshape(batch_size, -1)
  depth_mask = np.logical_and(depth > 0, depth < F32(depth_scale))
  filtered_feats = feats.reshape(batch_size, -1, channels)
  filtered_feats = filtered_feats * depth_mask[..., None].astype(np.int32)
  filtered_feats = filtered_feats.astype(F32)
  intrinsic = get_intrinsic_matrix(HFOV)
  xyz = xyz.reshape(batch_size, 4, -1)
  xyz = xyz * depth_mask[:, None, :].astype(F32)
  xyz = np.matmul(np.linalg.inv(intrinsic).astype(F32), xyz).astype(F32)
  if is_scalar_feat:
    filtered_feats = filtered_feats[..., 0]
  return xyz, filtered_feats


def project_to_feat(transformed_coords, feats, height, width, depth_scale,
                    input_void_class, output_void_class=0, return_debug: bool = False):
  """utils/point_cloud_utils.py:90-183.

  With return_debug=True additionally returns a dict with the per point flat
  index before / after the tolerance test, the per point depth and the raw
  (unclipped) z-buffer -- the quantities the parity bar calls "target-pixel
  indices" and "z-buffer".
  """
  feats = np.asarray(feats)
  if feats.ndim != 2 and feats.ndim != 3:
    raise ValueError('feats should have shape (N, M) or (N, M, C), got'
                     f' {feats.shape} instead.')
  is_scalar_feat = feats.ndim == 2
  if is_scalar_feat:
    feats = feats[..., None]
  channels = feats.shape[-1]
  transformed_coords = transformed_coords.astype(F32)
  feats = feats.astype(F32)
  batch_size = transformed_coords.shape[0]
  depth = transformed_coords[:, 2, :]
  view_coords = divide_no_nan(transformed_coords[:, 0:2, :], depth[:, None, :])
  fx = (view_coords[:, 0, :] + F32(1)) / F32(2) * F32(width)
  fy = (view_coords[:, 1, :] + F32(1)) / F32(2) * F32(height)
  col = cast_int32(fx)
  row = cast_int32(fy)
  valid = np.logical_and(np.logical_and(col >= 0, col < width),
                         np.logical_and(row >= 0, row < height))
  with np.errstate(invalid='ignore'):
    valid = np.logical_and(valid, depth > 0)
  valid_feats = np.all(feats != F32(input_void_class), axis=-1)
  valid = np.logical_and(valid, valid_feats)
  batch_offset = (np.arange(0, batch_size, dtype=np.int32)[:, None] * np.int32(width) * np.int32(height))
  with np.errstate(over='ignore'):
    flat = ((batch_offset + row * np.int32(width) + col) * valid.astype(np.int32)).astype(np.int32)
  flat = flat.reshape(-1)
  flat_depth = depth.reshape(-1)

  zinit = np.full((batch_size * height * width, 1), depth_scale, dtype=F32)
  scattered_depth = scatter_min(zinit[:, 0], flat, flat_depth)
  projected_depth = scattered_depth.reshape(batch_size, height, width)
  projected_depth = (np.clip(projected_depth, F32(0), F32(depth_scale)) / F32(depth_scale)).astype(F32)

  min_depth = scattered_depth[flat]
  with np.errstate(invalid='ignore'):
    keep = flat_depth < min_depth + F32(0.1)
  flat2 = flat * keep.astype(np.int32)

  flat_feats = feats.reshape(-1, channels)
  finit = np.full((batch_size * height * width, channels), output_void_class, dtype=F32)
  scattered_feat = scatter_max(finit, flat2, flat_feats)
  projected_feat = scattered_feat.reshape(batch_size, height, width, channels)
  if is_scalar_feat:
    projected_feat = projected_feat[..., 0]
  if return_debug:
    dbg = dict(flat=flat.reshape(batch_size, -1), flat_kept=flat2.reshape(batch_size, -1),
               valid=valid, depth=depth, zbuf=scattered_depth.reshape(batch_size, height, width))
    return projected_depth, projected_feat, dbg
  return projected_depth, projected_feat


# --------------------------------------------------------------------------
# Guidance assembly (models/models.py:282-293; trainers/gan_manager.py:484-494;
# utils/eval_metric.py:168-177)
# --------------------------------------------------------------------------
def guidance_from_projection(proj_depth: np.ndarray, proj_rgb: np.ndarray):
  """Returns (proj_image (N,H,W,3), proj_depth (N,H,W,1), proj_mask (N,H,W,1)) f32."""
  mask = np.logical_and(
      np.logical_and(proj_depth > 0, proj_depth < 1),
      np.all(proj_rgb != INVALID_RGB_VALUE, axis=-1)).astype(F32)[..., None]
  image = np.clip((proj_rgb / F32(255)).astype(F32), 0, 1).astype(F32)
  return image, proj_depth[..., None].astype(F32), mask


# --------------------------------------------------------------------------
# Caller glue: SE3DSModel memory (models/models.py:77-87,120-152,180-321)
# --------------------------------------------------------------------------
class MemoryState(NamedTuple):
  coords: np.ndarray      # (N, 4, M) f32
  feats: np.ndarray       # (N, M) u8 semantic   (reference keeps (N, M, 1) when empty)
  rgb_coords: np.ndarray  # (N, 4, M') f32
  rgb: np.ndarray         # (N, M', 3) int32


class SE3DSMemoryOracle:
  """The guidance half of models/models.py SE3DSModel (no generator)."""

  def __init__(self, height: int, depth_scale: float = DEPTH_SCALE, batch_size: int = 1,
               enforce_batch_one: bool = False):
    if enforce_batch_one and batch_size != 1:
      raise ValueError('Several methods do not support batch_size > 1.')
    self.batch_size = batch_size
    self.height = height
    self.width = 2 * height
    self.depth_scale = depth_scale
    self.reset_memory()

  def reset_memory(self):
    n = self.batch_size
    self._memory = MemoryState(
        coords=np.zeros((n, 4, 0), F32), feats=np.zeros((n, 0), np.uint8),
        rgb_coords=np.zeros((n, 4, 0), F32), rgb=np.zeros((n, 0, 3), np.int32))

  def get_memory_state(self) -> MemoryState:
    return MemoryState(*[a.copy() for a in self._memory])

  def _transform_position(self, xyz):
    xyz = np.asarray(xyz, F32)
    return np.stack([xyz[:, 0], xyz[:, 1], xyz[:, 2], np.zeros(self.batch_size, F32)], axis=1)

  def add_to_memory(self, pano_rgb, pano_semantic, pano_depth, position, mask_blurred=True):
    """models/models.py:180-245."""
    pano_rgb = np.asarray(pano_rgb).astype(np.int32)
    pano_semantic = np.asarray(pano_semantic).astype(np.uint8)
    if mask_blurred:
      pano_rgb = mask_pano(pano_rgb, masked_region_value=INVALID_RGB_VALUE)
    pos = self._transform_position(position)
    xyz1, feats = equirectangular_to_pointcloud(
        pano_semantic, pano_depth, INVALID_SEM_VALUE, self.depth_scale,
        interpolation_method='nearest')
    rgb_xyz1, rgb_feats = equirectangular_to_pointcloud(
        pano_rgb, pano_depth, INVALID_RGB_VALUE, self.depth_scale,
        interpolation_method='bilinear')
    xyz1 = xyz1 + pos[:, :, None]
    rgb_xyz1 = rgb_xyz1 + pos[:, :, None]
    feats_valid = np.any(feats != INVALID_SEM_VALUE, axis=(0, 2))
    rgb_valid = np.any(rgb_feats != INVALID_RGB_VALUE, axis=(0, 2))
    m = self._memory
    self._memory = MemoryState(
        coords=np.concatenate([m.coords, xyz1[:, :, feats_valid]], axis=2),
        feats=np.concatenate([m.feats, feats[:, feats_valid, 0]], axis=1),
        rgb_coords=np.concatenate([m.rgb_coords, rgb_xyz1[:, :, rgb_valid]], axis=2),
        rgb=np.concatenate([m.rgb, rgb_feats[:, rgb_valid, :].astype(np.int32)], axis=1))

  def project(self, position):
    """The guidance half of models/models.py:247-321.

    Returns dict(proj_image, proj_depth, proj_mask, proj_semantic, raw_rgb, raw_depth).
    """
    pos = self._transform_position(position)
    rel = self._memory.coords - pos[..., None]
    rel_rgb = self._memory.rgb_coords - pos[..., None]
    _, proj_sem = project_feats_to_equirectangular(
        self._memory.feats, rel, self.height, self.width, INVALID_SEM_VALUE, self.depth_scale)
    proj_depth, proj_rgb = project_feats_to_equirectangular(
        self._memory.rgb, rel_rgb, self.height, self.width, INVALID_RGB_VALUE, self.depth_scale)
    image, depth, mask = guidance_from_projection(proj_depth, proj_rgb)
    return dict(proj_image=image, proj_depth=depth, proj_mask=mask,
                proj_semantic=proj_sem.astype(np.uint8), raw_rgb=proj_rgb, raw_depth=proj_depth)


def reproject_trajectory(rgb, depth, src_pos, tgt_pos, depth_scale=DEPTH_SCALE,
                         unproject_void=INVALID_RGB_VALUE, project_void=INVALID_RGB_VALUE,
                         mask_first_frame: bool = True, mask_proportion: float = 0.125):
  """The rollout-loop form of the path (trainers/gan_manager.py:458-556,
  utils/eval_metric.py:144-240) for ONE target: S already-known source frames
  are unprojected, offset by their positions, concatenated unfiltered, offset by
  the target position and projected.

  rgb (N,S,H,W,3) integer in [0,255]; depth (N,S,H,W) f32; src_pos (N,S,3);
  tgt_pos (N,3).  Returns (proj_image, proj_depth, proj_mask, dbg).
  gan_manager uses unproject_void=0 / project_void=-1; eval_metric -1 / -1.
  """
  rgb = np.asarray(rgb)
  n, s, h, w, _ = rgb.shape
  coords, feats = [], []
  for k in range(s):
    frame = rgb[:, k].astype(np.int32)
    if mask_first_frame and k == 0:
      frame = mask_pano(frame, proportion=mask_proportion, masked_region_value=INVALID_RGB_VALUE)
    xyz1, f = equirectangular_to_pointcloud(frame, depth[:, k], unproject_void, depth_scale)
    p = np.concatenate([np.asarray(src_pos[:, k], F32), np.zeros((n, 1), F32)], axis=1)
    coords.append(xyz1 + p[:, :, None])
    feats.append(f)
  coords = np.concatenate(coords, axis=2)
  feats = np.concatenate(feats, axis=1)
  t = np.concatenate([np.asarray(tgt_pos, F32), np.zeros((n, 1), F32)], axis=1)
  rel = coords - t[:, :, None]
  pd, pf, dbg = project_feats_to_equirectangular(feats, rel, h, w, project_void, depth_scale,
                                                 return_debug=True)
  image, d, mask = guidance_from_projection(pd, pf)
  dbg['raw_rgb'] = pf
  return image, d, mask, dbg


# --------------------------------------------------------------------------
# inference/perturbation_utils.py
# --------------------------------------------------------------------------
def get_proportion_invalid_for_depth(position_offset, depth_image, distance_padding: float = 0.10):
  """inference/perturbation_utils.py:23-71 (float32 tensor scalars, python ints)."""
  po = np.asarray(position_offset, F32)
  depth_image = np.asarray(depth_image, F32)
  distance = np.sqrt(np.sum(po * po, dtype=F32), dtype=F32)
  height, width = depth_image.shape
  heading = np.arctan2(-po[0], -po[1])
  # `a + b * c % d` parses as a + ((b * c) % d); (2pi*{0,1}) % 2pi == 0 -> no-op.
  heading = heading + F32(np.fmod(F32(2 * math.pi) * F32(heading <= 0), F32(2 * math.pi)))
  if heading < 0:
    heading = F32(heading + F32(2 * math.pi))
  heading_proportion = F32(heading / F32(2 * math.pi))
  delta_xy = math.sqrt(float(F32(po[1] * po[1]) + F32(po[0] * po[0])))  # f32 squares, f64 sqrt
  elevation = np.arctan2(F32(delta_xy), -po[2])
  elevation = elevation + F32(np.fmod(F32(math.pi) * F32(elevation <= 0), F32(math.pi)))
  if elevation < 0:
    elevation = F32(elevation + F32(math.pi))
  elevation_proportion = F32(elevation / F32(math.pi))
  heading_start = int(F32(heading_proportion * F32(width)))
  elevation_start = int(F32(elevation_proportion * F32(height)))
  threshold_width = int(30 / 360 * width)
  threshold_height = int(60 / 180 * height)
  region = depth_image[
      max(0, elevation_start - threshold_height):min(height, elevation_start + threshold_height),
      max(0, heading_start - threshold_width):min(width, heading_start + threshold_width)]
  return np.mean(region * F32(DEPTH_SCALE) < F32(distance + F32(distance_padding)))


# --------------------------------------------------------------------------
# utils/pano_utils.py:92-114 (next-row function; pinned by the ray-table KAT)
# --------------------------------------------------------------------------
def equirectangular_pixel_rays(output_height: int) -> np.ndarray:
  output_width = int(F32(output_height) * 2)
  heading = tf_linspace(-math.pi, math.pi, output_width)
  pitch = tf_linspace(0.0, math.pi, output_height)
  heading, pitch = np.meshgrid(heading, pitch)
  xs = np.sin(pitch) * np.sin(heading)
  ys = -np.cos(pitch)
  zs = np.sin(pitch) * np.cos(heading)
  return np.stack([xs, ys, zs], axis=0).reshape(3, -1).astype(F32)


# --------------------------------------------------------------------------
# "Next" rows: bilinear resampling functions of utils/pano_utils.py
# --------------------------------------------------------------------------
def interpolate_bilinear(grid: np.ndarray, query_points: np.ndarray, indexing: str = 'ij') -> np.ndarray:
  """tensorflow_addons.image.interpolate_bilinear (tfa 0.16.1, dense_image_warp.py), float32.

  grid (B,H,W,C); query_points (B,N,2) in (y,x) order for 'ij', (x,y) for 'xy'.  Floors are clamped
  to [0, size-2], alphas to [0,1]; interp = a_y * (bottom - top) + top with top = a_x*(tr-tl)+tl.
  """
  grid = np.asarray(grid, F32)
  q = np.asarray(query_points, F32)
  b, h, w, c = grid.shape
  order = [0, 1] if indexing == 'ij' else [1, 0]
  floors, ceils, alphas = [], [], []
  for i, dim in enumerate(order):
    queries = q[..., dim]
    size = grid.shape[i + 1]
    floor = np.minimum(np.maximum(F32(0), np.floor(queries)), F32(size - 2)).astype(F32)
    int_floor = floor.astype(np.int32)
    floors.append(int_floor)
    ceils.append(int_floor + 1)
    alpha = np.minimum(np.maximum(F32(0), (queries - floor).astype(F32)), F32(1))
    alphas.append(alpha[..., None])
  bi = np.arange(b)[:, None]
  tl = grid[bi, floors[0], floors[1]]
  tr = grid[bi, floors[0], ceils[1]]
  bl = grid[bi, ceils[0], floors[1]]
  br = grid[bi, ceils[0], ceils[1]]
  top = alphas[1] * (tr - tl) + tl
  bottom = alphas[1] * (br - bl) + bl
  return (alphas[0] * (bottom - top) + top).astype(F32)


def get_world_to_image_transform(image_shape, fov, camera_intrinsics=None, rotations=None, rotation_matrix=None):
  """utils/pano_utils.py:26-89 (float32)."""
  if camera_intrinsics is None:
    height, width = F32(image_shape[0]), F32(image_shape[1])
    fov_y, fov_x = F32(fov[0]), F32(fov[1])
    fx = F32(0.5) * (width - F32(1.0)) / np.tan(fov_x / F32(2))
    fy = F32(0.5) * (height - F32(1.0)) / np.tan(fov_y / F32(2))
    camera_intrinsics = np.array([[fx, 0, F32(0.5) * (width - 1)], [0, fy, F32(0.5) * (height - 1)], [0, 0, 1]], F32)
  camera_intrinsics = np.asarray(camera_intrinsics, F32)
  if rotations is not None:
    rp, rh = F32(rotations[0]), F32(rotations[1])
    pitch = np.array([[1, 0, 0], [0, np.cos(-rp), -np.sin(-rp)], [0, np.sin(-rp), np.cos(-rp)]], F32)
    heading = np.array([[np.cos(-rh), 0, np.sin(-rh)], [0, 1, 0], [-np.sin(-rh), 0, np.cos(-rh)]], F32)
    extrinsics = (pitch @ heading).astype(F32)
  elif rotation_matrix is not None:
    extrinsics = np.asarray(rotation_matrix, F32)
  else:
    extrinsics = np.eye(3, dtype=F32)
  return (camera_intrinsics @ extrinsics).astype(F32)


def rotate_pano(pano: np.ndarray, matrix: np.ndarray, output_height: Optional[int] = None) -> np.ndarray:
  """utils/pano_utils.py:306-341."""
  pano = np.asarray(pano, F32)
  n, h, w, c = pano.shape
  if w != h * 2:
    raise ValueError('Pano width must be twice height.')
  oh = h if output_height is None else output_height
  ow = 2 * oh
  rays = equirectangular_pixel_rays(oh)
  rot = np.matmul(np.asarray(matrix, F32), rays[None]).astype(F32)
  x, y, z = rot[:, 0], rot[:, 1], rot[:, 2]
  with np.errstate(invalid='ignore'):
    pitch = np.arccos(-y)
  heading = np.arctan2(x, z)
  heading_pixels = (heading / F32(2 * math.pi) + F32(0.5)) * F32(w - 1)
  pitch_pixels = pitch / F32(math.pi) * F32(h - 1)
  coords = np.stack([pitch_pixels, heading_pixels], axis=-1).astype(F32)
  return interpolate_bilinear(pano, coords).reshape(n, oh, ow, c)


def project_perspective_image(image, fov, output_height, camera_intrinsics=None, rotations=None,
                              rotation_matrix=None, pad_mode='constant', pad_value=0.0, round_to_nearest=False):
  """utils/pano_utils.py:344-417."""
  assert pad_mode in {'reflect', 'constant', 'mean'}, 'Unsupported pad mode: %s' % pad_mode
  image = np.asarray(image, F32)[None]
  output_width = 2 * output_height
  world = equirectangular_pixel_rays(output_height)
  w2i = get_world_to_image_transform((image.shape[1], image.shape[2]), fov, camera_intrinsics=camera_intrinsics,
                                     rotations=rotations, rotation_matrix=rotation_matrix)
  ic = (w2i @ world).astype(F32).T
  xy, zs = ic[:, :2], ic[:, 2:]
  with np.errstate(divide='ignore', invalid='ignore'):
    ic = np.where(np.broadcast_to(zs > 0, xy.shape), xy / zs, -np.ones_like(xy)).astype(F32)
  if round_to_nearest:
    ic = np.round(ic)
  if pad_mode != 'reflect':
    cv = F32(image.mean(dtype=np.float64)) if pad_mode == 'mean' else F32(pad_value)
    image = np.pad(image, ((0, 0), (1, 1), (1, 1), (0, 0)), mode='constant', constant_values=cv)
    ic = ic + F32(1.0)
  out = interpolate_bilinear(image, ic[None], indexing='xy')
  return out.reshape(output_height, output_width, -1)


def get_perspective_from_equirectangular_image(image, camera_intrinsics, rotation_matrix, height, width):
  """utils/pano_utils.py:443-476."""
  image = np.asarray(image, F32)
  eq_h, eq_w, channels = image.shape
  x, y = np.meshgrid(np.arange(width), np.arange(height))
  xyz = np.stack([x, y, np.ones_like(x)], axis=-1).astype(F32)
  k_inv = np.linalg.inv(np.asarray(camera_intrinsics, F32)).astype(F32)
  xyz = ((xyz @ k_inv.T).astype(F32) @ np.asarray(rotation_matrix, F32)).astype(F32)
  norm = np.sqrt(np.sum(xyz * xyz, axis=-1, keepdims=True, dtype=F32))
  nrm = xyz / norm
  lon = np.arctan2(nrm[..., 0:1], nrm[..., 2:])
  lat = np.arcsin(nrm[..., 1:2])
  u = (lon / F32(2 * np.pi) + F32(0.5)) * F32(eq_w - 1)
  v = (lat / F32(np.pi) + F32(0.5)) * F32(eq_h - 1)
  uv = np.concatenate([u, v], axis=-1).astype(F32).reshape(-1, 2)
  out = interpolate_bilinear(image[None], uv[None], indexing='xy')
  return out.reshape(height, width, channels)


