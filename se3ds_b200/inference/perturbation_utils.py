"""Drop-in for the reference's inference/perturbation_utils.py, plus a batched form."""
from __future__ import annotations

import torch

from .. import _lib
from .. import constants


def get_proportion_invalid_for_depth_batch(position_offsets: torch.Tensor, depth_image: torch.Tensor,
                                           distance_padding: float = 0.10) -> torch.Tensor:
  """(P,3) offsets against one (H,W) depth map -> (P,) float32 proportions, one launch."""
  from ..utils.pano_utils import _as_tensor
  depth_image = _as_tensor(depth_image, 'depth_image').to(torch.float32).contiguous()
  if depth_image.dim() != 2:
    raise ValueError(f'depth_image should have shape (H, W), got {tuple(depth_image.shape)}')
  offs = _as_tensor(position_offsets, 'position_offsets').to(device=depth_image.device, dtype=torch.float32)
  offs = offs.reshape(-1, 3).contiguous()
  h, w = depth_image.shape
  out = torch.empty((offs.shape[0],), dtype=torch.float32, device=depth_image.device)
  _lib.check(_lib.load().se3ds_proportion_invalid(
      _lib.ptr(offs), offs.shape[0], _lib.ptr(depth_image), h, w, float(distance_padding),
      float(constants.DEPTH_SCALE), _lib.ptr(out), _lib.stream_handle(depth_image.device)))
  return out


def get_proportion_invalid_for_depth(position_offset: torch.Tensor, depth_image: torch.Tensor,
                                     distance_padding: float = 0.10) -> float:
  """Proportion of collided pixels when moving by position_offset (reference perturbation_utils.py:23-71).

  Args: position_offset (3,); depth_image (H,W) in [0,1]; distance_padding in metres.
  """
  return float(get_proportion_invalid_for_depth_batch(position_offset, depth_image, distance_padding)[0])
