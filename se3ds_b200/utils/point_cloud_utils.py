"""Drop-in for the hot-path functions of the reference's utils/point_cloud_utils.py."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .. import _lib
from .. import constants


def get_intrinsic_matrix(hfov: float) -> torch.Tensor:
  """Intrinsic matrix for a horizontal FOV (reference point_cloud_utils.py:23-29)."""
  t = 1 / np.tan(hfov / 2.)
  return torch.tensor([[t, 0., 0., 0.], [0., t, 0., 0.], [0., 0., 1, 0], [0., 0., 0, 1]], dtype=torch.float32)


def _project(coords, feats, height, width, depth_scale, input_void_class, output_void_class, mode,
             return_winner=False):
  from .pano_utils import _as_tensor, _canon_feats  # local import: pano_utils imports this module
  feats = _as_tensor(feats, 'feats', validate_only=True)
  if feats.dim() != 2 and feats.dim() != 3:
    raise ValueError('feats should have shape (N, M) or (N, M, C), got'
                     f' {tuple(feats.shape)} instead.')
  feats = _as_tensor(feats, 'feats')
  is_scalar_feat = feats.dim() == 2
  if is_scalar_feat:
    feats = feats[..., None]
  feats = _canon_feats(feats.contiguous())
  coords = _as_tensor(coords, 'coords').to(device=feats.device, dtype=torch.float32).contiguous()
  if coords.dim() != 3 or coords.shape[1] != 4:
    raise ValueError(f'coordinates should have shape (N, 4, M), got {tuple(coords.shape)} instead.')
  n, m, c = feats.shape
  if coords.shape[0] != n or coords.shape[2] != m:
    raise ValueError(f'coordinates {tuple(coords.shape)} do not match feats {tuple(feats.shape)}')
  depth = torch.empty((n, height, width), dtype=torch.float32, device=feats.device)
  out = torch.empty((n, height, width, c), dtype=torch.float32, device=feats.device)
  winner = torch.empty((n, height, width), dtype=torch.int32, device=feats.device) if return_winner else None
  ws = _lib.default_workspace(feats.device)
  _lib.check(_lib.load().se3ds_project_cloud(
      ws.handle, _lib.ptr(coords), _lib.ptr(feats), _lib.dtype_code(feats), n, m, c, int(height), int(width), mode,
      float(input_void_class), float(output_void_class), float(depth_scale), _lib.ptr(depth), _lib.ptr(out),
      _lib.ptr(winner), _lib.stream_handle(feats.device)))
  if is_scalar_feat:
    out = out[..., 0]
  return (depth, out, winner) if return_winner else (depth, out)


def project_to_feat(transformed_coords: torch.Tensor, feats: torch.Tensor, height: int, width: int,
                    depth_scale: float, input_void_class: float,
                    output_void_class: float = 0) -> Tuple[torch.Tensor, torch.Tensor]:
  """Splats features at pseudo-perspective coordinates (reference point_cloud_utils.py:90-183).

  Args: transformed_coords (N,4,M) of (x,y,z,1); feats (N,M) or (N,M,C); height; width;
  depth_scale; input_void_class; output_void_class.
  Returns: projected_depth (N,H,W) in [0,1]; projected_feat (N,H,W[,C]) float32.
  """
  return _project(transformed_coords, feats, height, width, depth_scale, input_void_class, output_void_class,
                  mode=1)


def get_filtered_coords_and_feats(feats: torch.Tensor, depth: torch.Tensor, depth_scale: float):
  """Legacy perspective unprojection (reference point_cloud_utils.py:32-87).

  Args: feats (N,H,W) or (N,H,W,C) integer features; depth (N,H,W) in [0,1]; depth_scale.
  Returns: xyz (N,4,H*W) float32; filtered feats (N,H*W[,C]) float32 (zero where depth is invalid).
  """
  from .pano_utils import _as_tensor, _canon_feats
  feats = _as_tensor(feats, 'feats', validate_only=True)
  if feats.dim() != 3 and feats.dim() != 4:
    raise ValueError('feats should have shape (N, H, W) or (N, H, W, C),'
                     f' got {tuple(feats.shape)} instead.')
  is_scalar_feat = feats.dim() == 3
  if is_scalar_feat:
    feats = feats[..., None]
  feats = _canon_feats(_as_tensor(feats, 'feats').contiguous())
  depth = _as_tensor(depth, 'depth').to(device=feats.device, dtype=torch.float32).contiguous()
  batch_size, height, width = depth.shape
  channels = feats.shape[-1]
  k_inv = torch.linalg.inv(get_intrinsic_matrix(constants.HFOV))  # host, 4x4, diagonal
  xyz = torch.empty((batch_size, 4, height * width), dtype=torch.float32, device=feats.device)
  out = torch.empty((batch_size, height * width, channels), dtype=torch.float32, device=feats.device)
  _lib.check(_lib.load().se3ds_filtered_coords_and_feats(
      _lib.ptr(feats), _lib.dtype_code(feats), _lib.ptr(depth), batch_size, height, width, channels, float(depth_scale),
      float(k_inv[0, 0]), float(k_inv[1, 1]), _lib.ptr(xyz), _lib.ptr(out), _lib.stream_handle(feats.device)))
  if is_scalar_feat:
    out = out[..., 0]
  return xyz, out
