"""Drop-in for the hot-path functions of the reference's utils/pano_utils.py.

Same names, positional order, keyword names, defaults and exception types; tensors are torch
CUDA tensors (or anything `torch.as_tensor` accepts, moved to the current CUDA device).  The
arithmetic runs in libse3ds_geom.so (sm_100a CUDA); see DESIGN.md for the canonical float32
definition of atan2 / acos / sin / cos that both the kernels and the oracle follow.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from .. import _lib
from . import point_cloud_utils

_UNSIGNED = (torch.uint8,) + tuple(getattr(torch, n) for n in ('uint16', 'uint32', 'uint64') if hasattr(torch, n))


def _as_tensor(x, name, validate_only=False):
  """torch tensor on the CUDA device; validate_only defers the move so that shape / dtype
  errors are raised first (and identically on a machine without a GPU)."""
  if not isinstance(x, torch.Tensor):
    x = torch.as_tensor(x)
  return x if validate_only else _lib.require_cuda(x, name)


def _canon_feats(feats: torch.Tensor) -> torch.Tensor:
  """Maps feature dtypes onto the three the C ABI speaks (u8, i32, f32)."""
  if feats.dtype in (torch.uint8, torch.int32, torch.float32):
    return feats
  if feats.dtype in (torch.int8, torch.int16, torch.int64, torch.bool):
    return feats.to(torch.int32)
  if feats.dtype in (torch.float16, torch.bfloat16, torch.float64):
    return feats.to(torch.float32)
  raise ValueError(f'unsupported feature dtype {feats.dtype}')


def _resize(images: torch.Tensor, size, method: str) -> torch.Tensor:
  """tf.image.resize (half-pixel centres) of an (N,H,W,C) device tensor: 'nearest' keeps the dtype,
  'bilinear' returns float32."""
  n, h, w, c = images.shape
  bilinear = method == 'bilinear'
  out = torch.empty((n, size[0], size[1], c), dtype=torch.float32 if bilinear else images.dtype, device=images.device)
  _lib.check(_lib.load().se3ds_resize(_lib.ptr(images), _lib.dtype_code(images), n, h, w, c, int(size[0]), int(size[1]),
                                      int(bilinear), _lib.ptr(out), _lib.stream_handle(images.device)))
  return out


def mask_pano(pano: torch.Tensor, proportion: float = 0.125, masked_region_value=0) -> torch.Tensor:
  """Masks the top and bottom `proportion` rows of a panorama (reference pano_utils.py:245-265).

  pano: (N, H, W, C).  Rows r < int(H*p) and r > H - int(H*p) are set to masked_region_value
  (row H - int(H*p) is kept, as in the reference).
  """
  pano = _as_tensor(pano, 'pano', validate_only=True)
  if pano.dim() != 4:
    raise ValueError(f'not enough values to unpack: pano should be (N, H, W, C), got {tuple(pano.shape)}')
  pano = _as_tensor(pano, 'pano')
  orig_dtype = pano.dtype
  work = _canon_feats(pano)
  n, h, w, c = work.shape
  out = torch.empty_like(work)
  _lib.check(_lib.load().se3ds_mask_pano(_lib.ptr(work), _lib.dtype_code(work), n, h, w, c, float(proportion),
                                         float(masked_region_value), _lib.ptr(out), _lib.stream_handle(work.device)))
  return out if out.dtype == orig_dtype else out.to(orig_dtype)


def equirectangular_to_pointcloud(feats: torch.Tensor, depth: torch.Tensor, void_class: float,
                                  depth_scale: float, size_mult: float = 1.0,
                                  interpolation_method: str = 'nearest') -> Tuple[torch.Tensor, torch.Tensor]:
  """Unprojects an equirectangular RGB-D pano (reference pano_utils.py:164-242).

  Args: feats (N,H,W) or (N,H,W,C); depth (N,H,W) in [0,1]; void_class; depth_scale.
  Returns: xyz1 (N,4,H*W) float32; filtered feats (N,H*W[,C]) -- input dtype for 'nearest',
  float32 for any other interpolation method (tf.image.resize returns float32).
  """
  feats = _as_tensor(feats, 'feats', validate_only=True)
  if feats.dim() != 3 and feats.dim() != 4:
    raise ValueError('feats should have shape (N, H, W) or (N, H, W, C),'
                     f' got {tuple(feats.shape)} instead.')
  if void_class < 0.0 and feats.dtype in _UNSIGNED:
    raise ValueError('feats datatype must be signed if the void class is negative')
  is_scalar_feat = feats.dim() == 3
  if is_scalar_feat:
    feats = feats[..., None]
  batch_size, height, width, channels = feats.shape
  assert width == 2 * height, 'Expected equirectangular input images'
  if interpolation_method not in ('nearest', 'bilinear'):
    raise NotImplementedError(f"interpolation_method '{interpolation_method}': only 'nearest' and 'bilinear' are built")
  feats = _as_tensor(feats, 'feats')
  orig_dtype = feats.dtype
  feats = _canon_feats(feats.contiguous())
  depth = _as_tensor(depth, 'depth').to(device=feats.device, dtype=torch.float32).contiguous()
  if tuple(depth.shape) != (batch_size, height, width):
    raise ValueError(f'depth should have shape {(batch_size, height, width)}, got {tuple(depth.shape)}')
  if size_mult != 1.0:  # reference pano_utils.py:203-208: depth 'nearest', features with the given method
    scaled = (int(height * size_mult), int(width * size_mult))
    depth = _resize(depth[..., None].contiguous(), scaled, 'nearest')[..., 0].contiguous()
    feats = _resize(feats, scaled, interpolation_method)
    height, width = scaled
  out_dtype = feats.dtype if interpolation_method == 'nearest' else torch.float32
  xyz1 = torch.empty((batch_size, 4, height * width), dtype=torch.float32, device=feats.device)
  out = torch.empty((batch_size, height * width, channels), dtype=out_dtype, device=feats.device)
  ws = _lib.default_workspace(feats.device)
  _lib.check(_lib.load().se3ds_unproject_equirect(
      ws.handle, _lib.ptr(feats), _lib.dtype_code(feats), _lib.ptr(depth), batch_size, height, width, channels,
      float(void_class), float(depth_scale), _lib.ptr(xyz1), _lib.ptr(out), _lib._DTYPES[out_dtype],
      _lib.stream_handle(feats.device)))
  if interpolation_method == 'nearest' and out.dtype != orig_dtype:
    out = out.to(orig_dtype)
  if is_scalar_feat:
    out = out[..., 0]
  return xyz1, out


def project_feats_to_equirectangular(feats: torch.Tensor, xyz1: torch.Tensor, height: int, width: int,
                                     void_class: float, depth_scale: float,
                                     return_winner: bool = False):
  """Projects point cloud features into an equirectangular image (reference pano_utils.py:117-161).

  Args: feats (N,M) or (N,M,C); xyz1 (N,4,M); height, width; void_class; depth_scale.
  Returns: depth (N,H,W) float32 in [0,1]; feats (N,H,W[,C]) float32 (output void class 0).
  With return_winner=True also the (N,H,W) int32 index of the nearest point per pixel (-1 none).
  """
  return point_cloud_utils._project(xyz1, feats, height, width, depth_scale, void_class, 0, mode=0,
                                    return_winner=return_winner)


# ------------------------------------------------------------------------------------------
# "Next" rows: the bilinear resampling functions of the reference's pano_utils.
# Each function is one fused CUDA kernel (query coordinates computed in-kernel, then the tfa-style
# bilinear sample); se3ds_interpolate_bilinear exposes the gather on its own.  Only the 3x3 camera
# matrices are host-side arithmetic, exactly as the reference builds them.
# ------------------------------------------------------------------------------------------
def interpolate_bilinear(grid: torch.Tensor, query_points: torch.Tensor, indexing: str = 'ij') -> torch.Tensor:
  """tensorflow_addons.image.interpolate_bilinear: grid (B,H,W,C), query_points (B,N,2) -> (B,N,C)."""
  if indexing not in ('ij', 'xy'):
    raise ValueError("Indexing mode must be 'ij' or 'xy'")
  grid = _as_tensor(grid, 'grid', validate_only=True)
  query_points = _as_tensor(query_points, 'query_points', validate_only=True)
  if grid.dim() != 4:
    raise ValueError('Grid must be 4D Tensor')
  if query_points.dim() != 3 or query_points.shape[2] != 2:
    raise ValueError('Query points must be 3 dimensional and size 2 in dim 2.')
  if grid.shape[1] < 2 or grid.shape[2] < 2:
    raise ValueError('Grid must be at least 2x2.')
  grid = _as_tensor(grid, 'grid').to(torch.float32).contiguous()
  query_points = _as_tensor(query_points, 'query_points').to(device=grid.device, dtype=torch.float32).contiguous()
  b, h, w, c = grid.shape
  nq = query_points.shape[1]
  out = torch.empty((b, nq, c), dtype=torch.float32, device=grid.device)
  _lib.check(_lib.load().se3ds_interpolate_bilinear(_lib.ptr(grid), _lib.ptr(query_points), b, h, w, c, nq,
                                                    int(indexing == 'xy'), _lib.ptr(out),
                                                    _lib.stream_handle(grid.device)))
  return out


def _mat9(m) -> "ctypes.Array":
  """Row-major 3x3 host matrix as a C float[9]."""
  import ctypes
  vals = [float(v) for v in torch.as_tensor(m, dtype=torch.float32).cpu().reshape(-1)]
  if len(vals) != 9:
    raise ValueError(f'expected a 3x3 matrix, got {len(vals)} values')
  return (ctypes.c_float * 9)(*vals)


def equirectangular_pixel_rays(output_height: int, device=None) -> torch.Tensor:
  """Unit-ball point of every equirectangular pixel, (3, H*2H) (reference pano_utils.py:92-114):
  x right, y down, z forward at the image centre."""
  device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
  if not torch.cuda.is_available():
    raise _lib.Se3dsError('no CUDA device: the geometric guidance path has no CPU fallback')
  output_height = int(output_height)
  out = torch.empty((3, output_height * 2 * output_height), dtype=torch.float32, device=device)
  _lib.check(_lib.load().se3ds_pixel_rays(output_height, _lib.ptr(out), _lib.stream_handle(device)))
  return out


def get_world_to_image_transform(image_shape, fov, camera_intrinsics: Optional[torch.Tensor] = None,
                                 rotations=None, rotation_matrix: Optional[torch.Tensor] = None) -> torch.Tensor:
  """3x3 world-to-image transform (reference pano_utils.py:26-89); tiny, evaluated on the host."""
  f32 = torch.float32
  if camera_intrinsics is None:
    height, width = float(image_shape[0]), float(image_shape[1])
    fov = torch.as_tensor(fov, dtype=f32).cpu()
    fov_y, fov_x = fov[0], fov[1]
    fx = 0.5 * (width - 1.0) / torch.tan(fov_x / 2)
    fy = 0.5 * (height - 1.0) / torch.tan(fov_y / 2)
    camera_intrinsics = torch.tensor([[fx, 0, 0.5 * (width - 1)], [0, fy, 0.5 * (height - 1)], [0., 0, 1]], dtype=f32)
  camera_intrinsics = torch.as_tensor(camera_intrinsics, dtype=f32).cpu()
  if rotations is not None:
    rot = torch.as_tensor(rotations, dtype=f32).cpu()
    rp, rh = rot[0], rot[1]
    pitch = torch.tensor([[1., 0, 0], [0, torch.cos(-rp), -torch.sin(-rp)], [0, torch.sin(-rp), torch.cos(-rp)]], dtype=f32)
    heading = torch.tensor([[torch.cos(-rh), 0, torch.sin(-rh)], [0., 1, 0], [-torch.sin(-rh), 0, torch.cos(-rh)]], dtype=f32)
    extrinsics = pitch @ heading
  elif rotation_matrix is not None:
    extrinsics = torch.as_tensor(rotation_matrix, dtype=f32).cpu()
  else:
    extrinsics = torch.eye(3, dtype=f32)
  return camera_intrinsics @ extrinsics


def crop_pano(pano: torch.Tensor, proportion: float = 0.125, method: str = 'bilinear',
              resize_to_original: bool = False) -> torch.Tensor:
  """Removes the top and bottom `proportion` rows (reference pano_utils.py:268-303)."""
  pano = _as_tensor(pano, 'pano', validate_only=True)
  if pano.dim() == 3:
    height = pano.shape[0]
  elif pano.dim() == 4:
    height = pano.shape[1]
  else:
    raise ValueError(f'pano should be of shape (N, H, W, C), got {tuple(pano.shape)} instead.')
  masked_height = int(height * proportion)
  cropped = pano[..., masked_height:height - masked_height, :, :].contiguous()
  if not resize_to_original:
    return cropped
  # tf.image.resize(..., antialias=True) back to (H, W): the crop is never taller than the original, so
  # this is always an enlargement, where the antialiased triangle kernel has scale 1 and its normalised
  # weights are the bilinear ones (width is unchanged: identity).  'nearest' ignores antialias.
  if method not in ('bilinear', 'nearest'):
    raise ValueError(f'Unsupported resize method: {method}')
  cropped = _as_tensor(cropped, 'pano')
  batched = cropped if cropped.dim() == 4 else cropped[None]
  width = batched.shape[2]
  resized = _resize(_canon_feats(batched).contiguous(), (height, width), method)
  resized = resized.to(cropped.dtype)          # tf.cast: truncation toward zero for integer dtypes
  return resized if cropped.dim() == 4 else resized[0]


def rotate_pano(pano: torch.Tensor, matrix: torch.Tensor, output_height: Optional[int] = None) -> torch.Tensor:
  """Rotates an equirectangular pano by (N,3,3) matrices (reference pano_utils.py:306-341)."""
  pano = _as_tensor(pano, 'pano', validate_only=True)
  if pano.dim() != 4:
    raise ValueError(f'pano should be (N, H, W, C), got {tuple(pano.shape)}')
  n, h, w, c = pano.shape
  if w != h * 2:
    raise ValueError('Pano width must be twice height.')
  pano = _as_tensor(pano, 'pano').to(torch.float32).contiguous()
  oh = h if output_height is None else int(output_height)
  matrix = _as_tensor(matrix, 'matrix').to(device=pano.device, dtype=torch.float32).reshape(n, 3, 3).contiguous()
  out = torch.empty((n, oh, 2 * oh, c), dtype=torch.float32, device=pano.device)
  _lib.check(_lib.load().se3ds_rotate_pano(_lib.ptr(pano), _lib.ptr(matrix), n, h, w, c, oh, _lib.ptr(out),
                                           _lib.stream_handle(pano.device)))
  return out


def project_perspective_image(image, fov, output_height, camera_intrinsics=None, rotations=None,
                              rotation_matrix=None, pad_mode='constant', pad_value=0.0, round_to_nearest=False):
  """Perspective (h,w,c) -> equirectangular (H,2H,c) (reference pano_utils.py:344-417)."""
  assert pad_mode in {'reflect', 'constant', 'mean'}, ('Unsupported pad mode: %s' % pad_mode)
  image = _as_tensor(image, 'image', validate_only=True)
  if image.dim() != 3:
    raise ValueError(f'image should be (height, width, channels), got {tuple(image.shape)}')
  image = _as_tensor(image, 'image').to(torch.float32).contiguous()
  h, w, c = image.shape
  w2i = get_world_to_image_transform((h, w), fov, camera_intrinsics=camera_intrinsics, rotations=rotations,
                                     rotation_matrix=rotation_matrix)
  pad = pad_mode != 'reflect'
  cv = float(image.mean()) if pad_mode == 'mean' else float(pad_value)
  out = torch.empty((int(output_height), 2 * int(output_height), c), dtype=torch.float32, device=image.device)
  _lib.check(_lib.load().se3ds_project_perspective_image(
      _lib.ptr(image), h, w, c, _mat9(w2i), int(output_height), int(pad), cv if pad else 0.0, int(bool(round_to_nearest)),
      _lib.ptr(out), _lib.stream_handle(image.device)))
  return out


def get_perspective_from_equirectangular_image(image, camera_intrinsics, rotation_matrix, height, width):
  """Equirectangular (He,We,C) -> perspective (height,width,C) (reference pano_utils.py:443-476)."""
  image = _as_tensor(image, 'image', validate_only=True)
  if image.dim() != 3:
    raise ValueError(f'image should be (H, W, C), got {tuple(image.shape)}')
  image = _as_tensor(image, 'image').to(torch.float32).contiguous()
  eq_height, eq_width, channels = image.shape
  k_inv_t = torch.linalg.inv(torch.as_tensor(camera_intrinsics, dtype=torch.float32).cpu()).t().contiguous()
  out = torch.empty((int(height), int(width), channels), dtype=torch.float32, device=image.device)
  _lib.check(_lib.load().se3ds_perspective_from_equirect(
      _lib.ptr(image), eq_height, eq_width, channels, _mat9(k_inv_t), _mat9(rotation_matrix), int(height), int(width),
      _lib.ptr(out), _lib.stream_handle(image.device)))
  return out
