"""Drop-in for the hot-path functions of the reference's utils/pano_utils.py.

Same names, positional order, keyword names, defaults and exception types; tensors are torch
CUDA tensors (or anything `torch.as_tensor` accepts, moved to the current CUDA device).  The
arithmetic runs in libse3ds_geom.so (sm_100a CUDA); see DESIGN.md for the canonical float32
definition of atan2 / acos / sin / cos that both the kernels and the oracle follow.
"""
from __future__ import annotations

from typing import Tuple

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
  if size_mult != 1.0:
    raise NotImplementedError('size_mult != 1.0 (resized clouds) is not on the accelerated path yet')
  feats = _as_tensor(feats, 'feats')
  orig_dtype = feats.dtype
  feats = _canon_feats(feats.contiguous())
  depth = _as_tensor(depth, 'depth').to(device=feats.device, dtype=torch.float32).contiguous()
  if tuple(depth.shape) != (batch_size, height, width):
    raise ValueError(f'depth should have shape {(batch_size, height, width)}, got {tuple(depth.shape)}')
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
