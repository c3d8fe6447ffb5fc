"""Fused guidance path: RGB-D panoramas + poses -> (proj_image, proj_depth, proj_mask).

`reproject` is the one-call form of what the reference spreads over
  pano_utils.mask_pano -> pano_utils.equirectangular_to_pointcloud -> `xyz1 += position` ->
  [compaction] -> tf.concat over frames -> `coords - position` ->
  pano_utils.project_feats_to_equirectangular -> guidance dict
(models/models.py:180-321, trainers/gan_manager.py:458-556, utils/eval_metric.py:144-240),
without ever materialising the point cloud.  `GuidanceMemory` mirrors the memory interface of
the reference's SE3DSModel (add_to_memory / __call__ / get, set, reset memory) on top of it,
keeping the S source frames (7 B per point) instead of a concatenated cloud (28 B per point).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from . import constants
from .utils import pano_utils


class Conventions(NamedTuple):
  """Void conventions of the three reference callers."""
  unproject_void: int
  project_void: int
  filter_void: bool

  def kwargs(self):
    """Keyword arguments for `reproject` / `prepare` / `reproject_host`."""
    return dict(unproject_void=self.unproject_void, project_void=self.project_void, filter_void=self.filter_void)


SE3DS_MODEL = Conventions(constants.INVALID_RGB_VALUE, constants.INVALID_RGB_VALUE, True)   # models/models.py
GAN_MANAGER = Conventions(0, constants.INVALID_RGB_VALUE, False)                            # trainers/gan_manager.py:476-548
EVAL_METRIC = Conventions(constants.INVALID_RGB_VALUE, constants.INVALID_RGB_VALUE, False)  # utils/eval_metric.py:163-236


def _prep(rgb, depth, src_pos, tgt_pos, to_device: bool):
  conv = (lambda t, n: _lib.require_cuda(torch.as_tensor(t), n)) if to_device else (lambda t, n: torch.as_tensor(t).contiguous())
  rgb = conv(rgb, 'rgb')
  if rgb.dim() == 4:
    rgb = rgb[:, None]
  if rgb.dim() != 5 or rgb.shape[-1] != 3:
    raise ValueError(f'rgb should have shape (N, S, H, W, 3), got {tuple(rgb.shape)} instead.')
  if rgb.dtype not in (torch.uint8, torch.int32):
    if rgb.dtype in (torch.int64, torch.int16, torch.int8):
      rgb = rgb.to(torch.int32)
    else:
      raise ValueError(f'rgb must be uint8 or int32 with values in [-1, 255], got {rgb.dtype}')
  n, s, h, w, _ = rgb.shape
  assert w == 2 * h, 'Expected equirectangular input images'
  depth = conv(depth, 'depth').to(torch.float32).reshape(n, s, h, w).contiguous()
  src_pos = conv(src_pos, 'src_pos').to(torch.float32).reshape(n, s, 3).contiguous()
  tgt_pos = conv(tgt_pos, 'tgt_pos').to(torch.float32)
  tgt_pos = tgt_pos.reshape(n, -1, 3).contiguous()
  return rgb.contiguous(), depth, src_pos, tgt_pos


def reproject(rgb, depth, src_pos, tgt_pos, depth_scale: float = constants.DEPTH_SCALE,
              mask_proportion: float = 0.125, mask_frames: int = 0,
              unproject_void: int = constants.INVALID_RGB_VALUE, project_void: int = constants.INVALID_RGB_VALUE,
              filter_void: bool = False, per_job_bin: bool = False, return_winner: bool = False,
              export_bin: bool = False, out: Optional[Dict[str, torch.Tensor]] = None,
              workspace: Optional[_lib.Workspace] = None, tgt_rot=None, key64: bool = False,
              raw_features: bool = False) -> Dict[str, torch.Tensor]:
  """Re-projects S source RGB-D panos per item onto P target poses per item.

  Args:
    rgb: (N,S,H,W,3) uint8, or int32 in [-1,255] ((N,H,W,3) is taken as S=1).
    depth: (N,S,H,W) float32 in [0,1].
    src_pos: (N,S,3) source positions;  tgt_pos: (N,P,3) or (N,3) target positions.
    mask_frames: the first `mask_frames` frames get mask_pano(., mask_proportion, -1).
    unproject_void / project_void / filter_void: see `Conventions`.
    per_job_bin: every (item,pose) job is its own reference call (batch 1).
    raw_features: proj_image holds the raw per-channel maxima instead of clip(x/255, 0, 1).
    tgt_rot: optional (N,P,3,3) rotations into the target camera frames (full SE(3) poses; the
      reference only translates, models/models.py:120-125): q = R (local + src - tgt).
  Returns dict with proj_image (J,H,W,3), proj_depth (J,H,W,1), proj_mask (J,H,W,1), and optionally
  winner (J,H,W) int32 and bin (5,) = (min depth, max R, G, B of the call's reject bin, depth of the
  owner pixel's own winner) with export_bin (see se3ds_apply_bin).
  """
  rgb, depth, src_pos, tgt_pos = _prep(rgb, depth, src_pos, tgt_pos, True)
  n, s, h, w, _ = rgb.shape
  p = tgt_pos.shape[1]
  j = n * p
  dev = rgb.device
  if out is None:
    out = {}
  def buf(name, shape, dtype=torch.float32):
    t = out.get(name)
    # a caller's buffer is written through its data pointer as a dense tensor: it must be one
    if t is None or tuple(t.shape) != shape or t.dtype != dtype or t.device != dev or not t.is_contiguous():
      t = out[name] = torch.empty(shape, dtype=dtype, device=dev)
    return t
  image = buf('proj_image', (j, h, w, 3))
  pdepth = buf('proj_depth', (j, h, w, 1))
  mask = buf('proj_mask', (j, h, w, 1))
  winner = buf('winner', (j, h, w), torch.int32) if return_winner else None
  binb = buf('bin', (5,)) if export_bin else None
  flags = ((_lib.FLAG_FILTER_VOID if filter_void else 0) | (_lib.FLAG_BIN_PER_JOB if per_job_bin else 0) |
           (_lib.FLAG_KEY64 if key64 else 0) | (_lib.FLAG_RAW_FEATURES if raw_features else 0))
  ws = workspace or _lib.default_workspace(dev)
  if tgt_rot is not None:
    tgt_rot = _lib.require_cuda(torch.as_tensor(tgt_rot), 'tgt_rot').to(device=dev, dtype=torch.float32)
    tgt_rot = tgt_rot.reshape(n, p, 3, 3).contiguous()
  _lib.check(_lib.load().se3ds_reproject_se3(
      ws.handle, _lib.ptr(rgb), _lib.dtype_code(rgb), _lib.ptr(depth), _lib.ptr(src_pos), _lib.ptr(tgt_pos),
      _lib.ptr(tgt_rot), n, s, p, h, w, float(depth_scale), float(mask_proportion), int(mask_frames),
      int(unproject_void), int(project_void), flags, _lib.ptr(image), _lib.ptr(pdepth), _lib.ptr(mask),
      _lib.ptr(winner), _lib.ptr(binb), _lib.stream_handle(dev)))
  return out


class PreparedReprojection:
  """A reprojection call with every argument pre-bound: `run()` is one C-ABI call and nothing
  else (no tensor bookkeeping on the host), suitable for CUDA-graph capture after one warm-up
  run.  The tensors passed to `prepare` are kept alive and read / written in place."""

  def __init__(self, tensors, out, args, ws):
    self.tensors, self.out, self._args, self.workspace = tensors, out, args, ws
    self._fn = _lib.load().se3ds_reproject
    self.device = out['proj_image'].device

  def run(self, stream=None):
    st = _lib.stream_handle(self.device) if stream is None else ctypes.c_void_p(stream)
    _lib.check(self._fn(*self._args, st))
    return self.out


def prepare(rgb, depth, src_pos, tgt_pos, depth_scale: float = constants.DEPTH_SCALE,
            mask_proportion: float = 0.125, mask_frames: int = 0,
            unproject_void: int = constants.INVALID_RGB_VALUE, project_void: int = constants.INVALID_RGB_VALUE,
            filter_void: bool = False, per_job_bin: bool = False, return_winner: bool = False,
            workspace: Optional[_lib.Workspace] = None, key64: bool = False,
            inputs_ready: bool = False) -> PreparedReprojection:
  """Same arguments as `reproject`; allocates the outputs once and returns a PreparedReprojection.
  inputs_ready=True promises that the input tensors are not written by whatever kernel runs right
  before each `run()` on the stream (SE3DS_FLAG_INPUTS_READY)."""
  rgb, depth, src_pos, tgt_pos = _prep(rgb, depth, src_pos, tgt_pos, True)
  n, s, h, w, _ = rgb.shape
  p = tgt_pos.shape[1]
  j = n * p
  dev = rgb.device
  out = dict(proj_image=torch.empty((j, h, w, 3), device=dev), proj_depth=torch.empty((j, h, w, 1), device=dev),
             proj_mask=torch.empty((j, h, w, 1), device=dev))
  if return_winner:
    out['winner'] = torch.empty((j, h, w), dtype=torch.int32, device=dev)
  flags = ((_lib.FLAG_FILTER_VOID if filter_void else 0) | (_lib.FLAG_BIN_PER_JOB if per_job_bin else 0) |
           (_lib.FLAG_KEY64 if key64 else 0) | (_lib.FLAG_INPUTS_READY if inputs_ready else 0))
  ws = workspace or _lib.default_workspace(dev)
  args = (ws.handle, _lib.ptr(rgb), _lib.dtype_code(rgb), _lib.ptr(depth), _lib.ptr(src_pos), _lib.ptr(tgt_pos),
          n, s, p, h, w, float(depth_scale), float(mask_proportion), int(mask_frames), int(unproject_void),
          int(project_void), flags, _lib.ptr(out['proj_image']), _lib.ptr(out['proj_depth']),
          _lib.ptr(out['proj_mask']), _lib.ptr(out.get('winner')), None)
  return PreparedReprojection((rgb, depth, src_pos, tgt_pos), out, args, ws)


def apply_bin(bin_values: torch.Tensor, out: Dict[str, torch.Tensor], depth_scale: float = constants.DEPTH_SCALE,
              raw_features: bool = False):
  """Applies a reduced reject bin to pixel (0,0) of job 0 of `out` (see se3ds_apply_bin).

  bin_values: (5,) = reduced (min depth, max R, G, B) followed by the owner call's own fifth value
  (depth of that pixel's own winner, as exported).  raw_features: `out` was produced with
  raw_features=True (raw per-channel maxima instead of clip(x / 255, 0, 1))."""
  dev = out['proj_image'].device
  assert bin_values.numel() == 5, 'bin is (min depth, max R, max G, max B, own winner depth)'
  _lib.check(_lib.load().se3ds_apply_bin(_lib.ptr(bin_values.contiguous()), float(depth_scale),
                                         _lib.FLAG_RAW_FEATURES if raw_features else 0,
                                         _lib.ptr(out['proj_image']), _lib.ptr(out['proj_depth']),
                                         _lib.ptr(out['proj_mask']), _lib.ptr(out.get('winner')),
                                         _lib.stream_handle(dev)))


def reproject_host(rgb, depth, src_pos, tgt_pos, depth_scale: float = constants.DEPTH_SCALE,
                   mask_proportion: float = 0.125, mask_frames: int = 0,
                   unproject_void: int = constants.INVALID_RGB_VALUE,
                   project_void: int = constants.INVALID_RGB_VALUE, filter_void: bool = False,
                   per_job_bin: bool = False, return_winner: bool = False,
                   out: Optional[Dict[str, torch.Tensor]] = None, device: int = 0,
                   workspace: Optional[_lib.Workspace] = None) -> Dict[str, torch.Tensor]:
  """`reproject` for HOST tensors (pinned memory recommended): host->device copies, the fused
  kernels and the device->host copies of the guidance tensors all happen inside the C ABI call
  (se3ds_reproject_host), which returns when the host outputs are complete."""
  rgb, depth, src_pos, tgt_pos = _prep(rgb, depth, src_pos, tgt_pos, False)
  for t in (rgb, depth, src_pos, tgt_pos):
    if t.is_cuda:
      raise ValueError('reproject_host takes host tensors; use reproject for device tensors')
  n, s, h, w, _ = rgb.shape
  p = tgt_pos.shape[1]
  j = n * p
  if out is None:
    out = {}
  pin = torch.cuda.is_available()
  def buf(name, shape, dtype=torch.float32):
    t = out.get(name)
    if t is None or tuple(t.shape) != shape or t.dtype != dtype or t.is_cuda or not t.is_contiguous():
      t = out[name] = torch.empty(shape, dtype=dtype, pin_memory=pin)
    return t
  image = buf('proj_image', (j, h, w, 3))
  pdepth = buf('proj_depth', (j, h, w, 1))
  mask = buf('proj_mask', (j, h, w, 1))
  winner = buf('winner', (j, h, w), torch.int32) if return_winner else None
  flags = (_lib.FLAG_FILTER_VOID if filter_void else 0) | (_lib.FLAG_BIN_PER_JOB if per_job_bin else 0)
  ws = workspace or _lib.default_workspace(torch.device('cuda', device), role='host')
  _lib.check(_lib.load().se3ds_reproject_host(
      ws.handle, _lib.ptr(rgb), _lib.dtype_code(rgb), _lib.ptr(depth), _lib.ptr(src_pos), _lib.ptr(tgt_pos),
      n, s, p, h, w, float(depth_scale), float(mask_proportion), int(mask_frames), int(unproject_void),
      int(project_void), flags, _lib.ptr(image), _lib.ptr(pdepth), _lib.ptr(mask), _lib.ptr(winner)))
  return out


class MemoryState(NamedTuple):
  """Frame-ring memory: the S observations instead of the reference's concatenated cloud
  (models/models.py:77-87 keeps coords (N,4,M) / feats / rgb_coords / rgb)."""
  rgb: List[torch.Tensor]        # each (N,H,W,3) uint8 or int32
  semantic: List[torch.Tensor]   # each (N,H,W,1) uint8
  depth: List[torch.Tensor]      # each (N,H,W) float32
  position: List[torch.Tensor]   # each (N,3) float32
  masked: List[bool]


class GuidanceMemory(object):
  """The guidance half of the reference's SE3DSModel (models/models.py:90-321).

  add_to_memory(...) stores an observation; __call__(position) returns the generator inputs
  `proj_image`, `proj_depth`, `proj_mask`, `blurred_mask` (+ `proj_semantic`) for a target
  position.  The generator itself (models/models.py:323-366) is out of scope.
  """

  def __init__(self, image_height: int, depth_scale: float = constants.DEPTH_SCALE, batch_size: int = 1,
               project_semantic: bool = False):
    if batch_size != 1:
      raise ValueError('Several methods do not support batch_size > 1.')
    self.batch_size = batch_size
    self.height = image_height
    self.width = image_height * 2
    self.depth_scale = depth_scale
    self.project_semantic = project_semantic
    self.reset_memory()

  def _check_batch_size(self, input_batch_size):
    if input_batch_size != self.batch_size:
      raise ValueError('Input batch size is not suitable. Expected '
                       f'{self.batch_size}, got {input_batch_size} instead.')

  def reset_memory(self):
    self._memory = MemoryState([], [], [], [], [])

  def get_memory_state(self) -> MemoryState:
    m = self._memory
    return MemoryState([t.clone() for t in m.rgb], [t.clone() for t in m.semantic],
                       [t.clone() for t in m.depth], [t.clone() for t in m.position], list(m.masked))

  def set_memory_state(self, state: MemoryState):
    self._memory = MemoryState([t.clone() for t in state.rgb], [t.clone() for t in state.semantic],
                               [t.clone() for t in state.depth], [t.clone() for t in state.position],
                               list(state.masked))

  def add_to_memory(self, pano_rgb, pano_semantic, pano_depth, position, mask_blurred=True):
    """models/models.py:180-245 (the cloud is not built; the frame is kept)."""
    pano_semantic = _lib.require_cuda(torch.as_tensor(pano_semantic), 'pano_semantic')
    self._check_batch_size(pano_semantic.shape[0])
    pano_rgb = _lib.require_cuda(torch.as_tensor(pano_rgb), 'pano_rgb')
    assert pano_rgb.dtype in (torch.uint8, torch.int32)
    assert pano_semantic.dtype in (torch.uint8, torch.int32)
    m = self._memory
    m.rgb.append(pano_rgb)
    m.semantic.append(pano_semantic.to(torch.uint8))
    m.depth.append(_lib.require_cuda(torch.as_tensor(pano_depth), 'pano_depth').to(torch.float32))
    m.position.append(_lib.require_cuda(torch.as_tensor(position), 'position').to(torch.float32).reshape(-1, 3))
    m.masked.append(bool(mask_blurred))

  def __call__(self, position) -> Dict[str, torch.Tensor]:
    """The guidance half of models/models.py:247-321 for one (N,3) target position."""
    position = _lib.require_cuda(torch.as_tensor(position), 'position').to(torch.float32).reshape(-1, 3)
    self._check_batch_size(position.shape[0])
    m = self._memory
    if not m.rgb:
      raise ValueError('memory is empty: call add_to_memory first')
    # frames that want the blurred rows masked go first (mask_frames is a prefix count)
    order = sorted(range(len(m.rgb)), key=lambda i: not m.masked[i])
    dt = torch.int32 if any(t.dtype == torch.int32 for t in m.rgb) else torch.uint8
    rgb = torch.stack([m.rgb[i].to(dt) for i in order], dim=1)
    depth = torch.stack([m.depth[i] for i in order], dim=1)
    src = torch.stack([m.position[i] for i in order], dim=1)
    out = reproject(rgb, depth, src, position, self.depth_scale, mask_frames=sum(m.masked),
                    unproject_void=SE3DS_MODEL.unproject_void, project_void=SE3DS_MODEL.project_void,
                    filter_void=True, per_job_bin=True)
    out['blurred_mask'] = torch.zeros_like(out['proj_mask'])
    if self.project_semantic:
      out['proj_semantic'] = self._project_semantic(position)
    return out

  def rgb_cloud(self):
    """Materialises the RGB memory as the reference keeps it: rgb_coords (N,4,M'), rgb (N,M',3) int32
    after masking, unprojection, offset and compaction (models/models.py:211-245)."""
    m = self._memory
    coords, feats = [], []
    for rgb, depth, pos, masked in zip(m.rgb, m.depth, m.position, m.masked):
      rgb = rgb.to(torch.int32)
      if masked:
        rgb = pano_utils.mask_pano(rgb, masked_region_value=constants.INVALID_RGB_VALUE)
      xyz1, f = pano_utils.equirectangular_to_pointcloud(rgb, depth, constants.INVALID_RGB_VALUE, self.depth_scale)
      xyz1 = xyz1 + torch.cat([pos, torch.zeros_like(pos[:, :1])], dim=1)[:, :, None]
      valid = (f != constants.INVALID_RGB_VALUE).any(dim=0).any(dim=-1)
      coords.append(xyz1[:, :, valid])
      feats.append(f[:, valid])
    if not coords:
      dev = torch.device('cuda', torch.cuda.current_device())
      return torch.zeros((self.batch_size, 4, 0), device=dev), torch.zeros((self.batch_size, 0, 3), dtype=torch.int32, device=dev)
    return torch.cat(coords, dim=2), torch.cat(feats, dim=1)

  def write_memory_as_pointcloud(self, filename):
    """Writes memory at batch position 0 to an ASCII .ply file, byte for byte in the reference's
    format (models/models.py:154-178)."""
    coords, rgb = self.rgb_cloud()
    xyz_pts = coords[0, 0:3].cpu().numpy().T
    rgb_pts = rgb[0].cpu().numpy()
    with open(filename, 'w') as fp:
      fp.write('ply\n')
      fp.write('format ascii 1.0 \n')
      fp.write('element vertex %d\n' % xyz_pts.shape[0])
      fp.write('property float x\n')
      fp.write('property float y\n')
      fp.write('property float z\n')
      fp.write('property uchar red\n')
      fp.write('property uchar green\n')
      fp.write('property uchar blue\n')
      fp.write('end_header\n')
      for i in range(xyz_pts.shape[0]):
        fp.write('{} {} {} {} {} {} \n'.format(xyz_pts[i, 0], xyz_pts[i, 1], xyz_pts[i, 2], rgb_pts[i, 0],
                                               rgb_pts[i, 1], rgb_pts[i, 2]))

  def _project_semantic(self, position):
    """models/models.py:217-219,229-231,276-278 through the fused kernels: the class ids are
    replicated into the three feature channels (void class 0 on both sides, compaction on, no row
    mask), so validity, the near-min set and the per-channel maximum are those of the scalar
    feature, and channel 0 of the raw features is proj_semantic."""
    m = self._memory
    sem = torch.stack([t.reshape(t.shape[0], self.height, self.width) for t in m.semantic], dim=1)  # (N,S,H,W)
    sem3 = sem[..., None].expand(-1, -1, -1, -1, 3).contiguous()
    depth = torch.stack(m.depth, dim=1)
    src = torch.stack(m.position, dim=1)
    out = reproject(sem3, depth, src, position, self.depth_scale, mask_frames=0,
                    unproject_void=constants.INVALID_SEM_VALUE, project_void=constants.INVALID_SEM_VALUE,
                    filter_void=True, per_job_bin=True, raw_features=True)
    return out['proj_image'][..., 0].to(torch.uint8)
