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
from typing import Callable, Dict, List, NamedTuple, Optional

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
  if rgb.dtype == torch.int32 and rgb.numel():
    # The fused kernels reduce the colours in float16 (one 8-byte vector reduction per point), which holds the
    # integers of [-2048, 2048] exactly -- every value the reference produces lies in [-1, 255].  Anything else
    # would be rounded silently, so it is refused here (one reduction and a sync, int32 inputs only); arbitrary
    # float / int features go through pano_utils.project_feats_to_equirectangular, which keeps float32.
    lo, hi = torch.aminmax(rgb)
    if int(lo) < -2048 or int(hi) > 2048:
      raise ValueError(f'int32 rgb values must lie in [-2048, 2048] for the fused path, got [{int(lo)}, {int(hi)}]')
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
              raw_features: bool = False, frames: Optional[int] = None, compact: bool = False) -> Dict[str, torch.Tensor]:
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
    frames: only the first `frames` of the S frame slots are in use (a partly filled frame ring, see
      `FrameRing`); the tensors are read in place, nothing is sliced or copied.
    compact: the compact guidance format (SE3DS_FLAG_COMPACT_OUT): the dict holds proj_rgb_u8 (J,H,W,3) uint8
      and proj_depth instead of the three float32 tensors -- 7 instead of 20 bytes per pixel for copies and
      collectives; `expand_guidance` restores proj_image / proj_mask bit for bit.
  Returns dict with proj_image (J,H,W,3), proj_depth (J,H,W,1), proj_mask (J,H,W,1), and optionally
  winner (J,H,W) int32 and bin (5,) = (min depth, max R, G, B of the call's reject bin, depth of the
  owner pixel's own winner) with export_bin (see se3ds_apply_bin).
  """
  rgb, depth, src_pos, tgt_pos = _prep(rgb, depth, src_pos, tgt_pos, True)
  n, s_cap, h, w, _ = rgb.shape
  s = s_cap if frames is None else int(frames)
  if not 0 < s <= s_cap:
    raise ValueError(f'frames must be in [1, {s_cap}], got {frames}')
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
  if compact:
    image, mask = buf('proj_rgb_u8', (j, h, w, 3), torch.uint8), None
  else:
    image, mask = buf('proj_image', (j, h, w, 3)), buf('proj_mask', (j, h, w, 1))
  pdepth = buf('proj_depth', (j, h, w, 1))
  winner = buf('winner', (j, h, w), torch.int32) if return_winner else None
  binb = buf('bin', (5,)) if export_bin else None
  flags = ((_lib.FLAG_FILTER_VOID if filter_void else 0) | (_lib.FLAG_BIN_PER_JOB if per_job_bin else 0) |
           (_lib.FLAG_KEY64 if key64 else 0) | (_lib.FLAG_RAW_FEATURES if raw_features else 0) |
           (_lib.FLAG_COMPACT_OUT if compact else 0))
  ws = workspace or _lib.default_workspace(dev)
  if tgt_rot is not None:
    tgt_rot = _lib.require_cuda(torch.as_tensor(tgt_rot), 'tgt_rot').to(device=dev, dtype=torch.float32)
    tgt_rot = tgt_rot.reshape(n, p, 3, 3).contiguous()
  _lib.check(_lib.load().se3ds_reproject_ring(
      ws.handle, _lib.ptr(rgb), _lib.dtype_code(rgb), _lib.ptr(depth), _lib.ptr(src_pos), _lib.ptr(tgt_pos),
      _lib.ptr(tgt_rot), n, s, s_cap, p, h, w, float(depth_scale), float(mask_proportion), int(mask_frames),
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
    self.device = out['proj_depth'].device

  def run(self, stream=None):
    st = _lib.stream_handle(self.device) if stream is None else ctypes.c_void_p(stream)
    _lib.check(self._fn(*self._args, st))
    return self.out

  def redirect_outputs(self, image_ptr: int, depth_ptr: int):
    """The resolve kernel stores the (compact) colour and depth planes at these device addresses instead of
    `out`'s tensors -- e.g. the NVSwitch multicast mapping of a symmetric buffer (parallel.ShardedReprojection,
    wire='multicast': one store reaches every rank).  The addresses are only ever written."""
    a = list(self._args)
    a[17], a[18] = ctypes.c_void_p(image_ptr), ctypes.c_void_p(depth_ptr)
    self._args = tuple(a)


def prepare(rgb, depth, src_pos, tgt_pos, depth_scale: float = constants.DEPTH_SCALE,
            mask_proportion: float = 0.125, mask_frames: int = 0,
            unproject_void: int = constants.INVALID_RGB_VALUE, project_void: int = constants.INVALID_RGB_VALUE,
            filter_void: bool = False, per_job_bin: bool = False, return_winner: bool = False,
            workspace: Optional[_lib.Workspace] = None, key64: bool = False,
            inputs_ready: bool = False, compact: bool = False, out: Optional[Dict[str, torch.Tensor]] = None,
            bin_out: Optional[torch.Tensor] = None) -> PreparedReprojection:
  """Same arguments as `reproject`; allocates the outputs once (or takes the dense, contiguous tensors of `out`)
  and returns a PreparedReprojection.  inputs_ready=True promises that the input tensors are not written by
  whatever kernel runs right before each `run()` on the stream (SE3DS_FLAG_INPUTS_READY).  bin_out: a (5,)
  float32 device tensor that receives the call's reject bin instead of its owner pixel (see `reproject`)."""
  rgb, depth, src_pos, tgt_pos = _prep(rgb, depth, src_pos, tgt_pos, True)
  n, s, h, w, _ = rgb.shape
  p = tgt_pos.shape[1]
  j = n * p
  dev = rgb.device
  given = out or {}
  def buf(name, shape, dtype=torch.float32):
    t = given.get(name)
    if t is not None:
      if tuple(t.shape) != shape or t.dtype != dtype or t.device != dev or not t.is_contiguous():
        raise ValueError(f'out[{name!r}] must be a contiguous {dtype} tensor of shape {shape} on {dev}')
      return t
    return torch.empty(shape, dtype=dtype, device=dev)
  if compact:
    out = dict(proj_rgb_u8=buf('proj_rgb_u8', (j, h, w, 3), torch.uint8), proj_depth=buf('proj_depth', (j, h, w, 1)))
  else:
    out = dict(proj_image=buf('proj_image', (j, h, w, 3)), proj_depth=buf('proj_depth', (j, h, w, 1)),
               proj_mask=buf('proj_mask', (j, h, w, 1)))
  if return_winner:
    out['winner'] = buf('winner', (j, h, w), torch.int32)
  if bin_out is not None:
    out['bin'] = bin_out
  flags = ((_lib.FLAG_FILTER_VOID if filter_void else 0) | (_lib.FLAG_BIN_PER_JOB if per_job_bin else 0) |
           (_lib.FLAG_KEY64 if key64 else 0) | (_lib.FLAG_INPUTS_READY if inputs_ready else 0) |
           (_lib.FLAG_COMPACT_OUT if compact else 0))
  ws = workspace or _lib.default_workspace(dev)
  args = (ws.handle, _lib.ptr(rgb), _lib.dtype_code(rgb), _lib.ptr(depth), _lib.ptr(src_pos), _lib.ptr(tgt_pos),
          n, s, p, h, w, float(depth_scale), float(mask_proportion), int(mask_frames), int(unproject_void),
          int(project_void), flags, _lib.ptr(out['proj_rgb_u8'] if compact else out['proj_image']), _lib.ptr(out['proj_depth']),
          _lib.ptr(out.get('proj_mask')), _lib.ptr(out.get('winner')), _lib.ptr(bin_out))
  return PreparedReprojection((rgb, depth, src_pos, tgt_pos), out, args, ws)


def expand_guidance(out: Dict[str, torch.Tensor], job_map: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
  """Compact guidance (proj_rgb_u8 + proj_depth) -> adds the float32 proj_image (J,H,W,3) and proj_mask
  (J,H,W,1) of the reference contract, bit-identical to a non-compact `reproject` (se3ds_expand_guidance).
  job_map (J,) int32: source job s becomes job job_map[s] of the outputs (proj_depth is re-ordered too)."""
  rgb8, depth = out['proj_rgb_u8'].contiguous(), out['proj_depth'].contiguous()
  j, h, w, _ = rgb8.shape
  dev = rgb8.device
  image = torch.empty((j, h, w, 3), dtype=torch.float32, device=dev)
  mask = torch.empty((j, h, w, 1), dtype=torch.float32, device=dev)
  depth_out = None
  if job_map is not None:
    job_map = _lib.require_cuda(torch.as_tensor(job_map), 'job_map').to(device=dev, dtype=torch.int32).contiguous()
    depth_out = torch.empty_like(depth)
  _lib.check(_lib.load().se3ds_expand_guidance(_lib.ptr(rgb8), _lib.ptr(depth), j, h * w, _lib.ptr(job_map), _lib.ptr(image),
                                               _lib.ptr(depth_out), _lib.ptr(mask), _lib.stream_handle(dev)))
  out['proj_image'], out['proj_mask'] = image, mask
  if depth_out is not None:  # everything in the result is in job order: the compact colours are dropped, not re-ordered
    out['proj_depth'] = depth_out
    out.pop('proj_rgb_u8')
    if 'winner' in out:
      out['winner'] = torch.empty_like(out['winner']).index_copy_(0, job_map.long(), out['winner'])
  return out


def apply_bin(bin_values: torch.Tensor, out: Dict[str, torch.Tensor], depth_scale: float = constants.DEPTH_SCALE,
              raw_features: bool = False):
  """Applies a reduced reject bin to pixel (0,0) of job 0 of `out` (see se3ds_apply_bin).

  bin_values: (5,) = reduced (min depth, max R, G, B) followed by the owner call's own fifth value
  (depth of that pixel's own winner, as exported).  raw_features: `out` was produced with
  raw_features=True (raw per-channel maxima instead of clip(x / 255, 0, 1))."""
  dev = out['proj_depth'].device
  assert bin_values.numel() == 5, 'bin is (min depth, max R, max G, max B, own winner depth)'
  compact = 'proj_rgb_u8' in out and 'proj_image' not in out
  flags = (_lib.FLAG_RAW_FEATURES if raw_features else 0) | (_lib.FLAG_COMPACT_OUT if compact else 0)
  _lib.check(_lib.load().se3ds_apply_bin(_lib.ptr(bin_values.contiguous()), float(depth_scale), flags,
                                         _lib.ptr(out['proj_rgb_u8'] if compact else out['proj_image']), _lib.ptr(out['proj_depth']),
                                         _lib.ptr(out.get('proj_mask')), _lib.ptr(out.get('winner')),
                                         _lib.stream_handle(dev)))


def reproject_host(rgb, depth, src_pos, tgt_pos, depth_scale: float = constants.DEPTH_SCALE,
                   mask_proportion: float = 0.125, mask_frames: int = 0,
                   unproject_void: int = constants.INVALID_RGB_VALUE,
                   project_void: int = constants.INVALID_RGB_VALUE, filter_void: bool = False,
                   per_job_bin: bool = False, return_winner: bool = False,
                   out: Optional[Dict[str, torch.Tensor]] = None, device: int = 0,
                   workspace: Optional[_lib.Workspace] = None, compact: bool = False,
                   wait: bool = True) -> Dict[str, torch.Tensor]:
  """`reproject` for HOST tensors (pinned memory recommended): host->device copies, the fused
  kernels and the device->host copies of the guidance tensors all happen inside the C ABI call
  (se3ds_reproject_host), which returns when the host outputs are complete.  wait=False
  (SE3DS_FLAG_HOST_ASYNC) returns once the work is enqueued: inputs and outputs must be left alone until
  `workspace.host_wait()`; see HostReprojector for the double-buffered loop built on it."""
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
  if compact:  # 7 instead of 20 bytes per pixel come back over PCIe
    image, mask = buf('proj_rgb_u8', (j, h, w, 3), torch.uint8), None
  else:
    image, mask = buf('proj_image', (j, h, w, 3)), buf('proj_mask', (j, h, w, 1))
  pdepth = buf('proj_depth', (j, h, w, 1))
  winner = buf('winner', (j, h, w), torch.int32) if return_winner else None
  flags = ((_lib.FLAG_FILTER_VOID if filter_void else 0) | (_lib.FLAG_BIN_PER_JOB if per_job_bin else 0) |
           (_lib.FLAG_COMPACT_OUT if compact else 0) | (0 if wait else _lib.FLAG_HOST_ASYNC))
  if not wait and workspace is None:
    raise ValueError('wait=False needs an explicit workspace (the one to call host_wait() on)')
  ws = workspace or _lib.default_workspace(torch.device('cuda', device), role='host')
  _lib.check(_lib.load().se3ds_reproject_host(
      ws.handle, _lib.ptr(rgb), _lib.dtype_code(rgb), _lib.ptr(depth), _lib.ptr(src_pos), _lib.ptr(tgt_pos),
      n, s, p, h, w, float(depth_scale), float(mask_proportion), int(mask_frames), int(unproject_void),
      int(project_void), flags, _lib.ptr(image), _lib.ptr(pdepth), _lib.ptr(mask), _lib.ptr(winner)))
  return out


class HostReprojector(object):
  """Double-buffered `reproject_host`: `depth` workspaces (each with its own streams and staging buffers)
  are used in turn, so batch i+1 travels to the device while batch i is computed and travels back -- the
  host link is busy in both directions.  The reference has no counterpart (its tensors live wherever
  TensorFlow puts them); this is what a server that feeds host batches calls.

    pipe = HostReprojector(device=0, depth=2, mask_frames=1, compact=True)
    for batch in batches:
      done = pipe.submit(*batch)      # None while the pipeline fills, else the oldest finished result
      if done is not None: consume(done)
    for done in pipe.flush(): consume(done)

  The tensors handed to submit() must stay untouched until their result comes back.  A returned result
  (pinned host tensors, see reproject_host) stays valid until the next submit(): there is one more set of
  output buffers than workspaces."""

  def __init__(self, device: int = 0, depth: int = 2, **kwargs):
    if depth < 1:
      raise ValueError('depth must be >= 1')
    self._kw = dict(kwargs, device=device)
    self._ws = [_lib.Workspace(device) for _ in range(depth)]
    self._out = [dict() for _ in range(depth + 1)]
    self._calls = 0
    self._pending = []  # (workspace slot, output slot, inputs kept alive) in submission order

  def submit(self, rgb, depth, src_pos, tgt_pos):
    slot, oslot = self._calls % len(self._ws), self._calls % len(self._out)
    done = self._finish() if len(self._pending) == len(self._ws) else None  # the oldest call owns this workspace
    reproject_host(rgb, depth, src_pos, tgt_pos, out=self._out[oslot], workspace=self._ws[slot], wait=False, **self._kw)
    self._pending.append((slot, oslot, (rgb, depth, src_pos, tgt_pos)))
    self._calls += 1
    return done

  def _finish(self):
    slot, oslot, _ = self._pending.pop(0)
    self._ws[slot].host_wait()
    return self._out[oslot]

  def flush(self):
    """Completes everything in flight; -> the results in submission order (all valid until the next submit())."""
    res = []
    while self._pending:
      res.append(self._finish())
    return res

  def close(self):
    """Completes what is in flight and releases the workspaces (device staging buffers, streams)."""
    self.flush()
    for ws in self._ws:
      ws.close()
    self._ws = []


class FrameRing(object):
  """The memory of a trajectory as the frames that were observed: rgb (N,cap,H,W,3) uint8 / int32,
  depth (N,cap,H,W) float32, position (N,cap,3) float32, `count` slots in use.

  The reference keeps the concatenated cloud and re-concatenates it for every new frame
  (models/models.py:239-245, trainers/gan_manager.py:550-551, utils/eval_metric.py:238-239: O(M) per
  step, 28 B per point); here a frame is appended by writing its slot in place (7 - 16 B per point) and
  `reproject(..., frames=count)` reads the ring where it lies.  The capacity doubles when it runs out.
  Frames whose blurred rows are to be masked (`mask_pano`) are kept in front: `masked` is the prefix
  count `reproject(mask_frames=...)` expects (the order of the frames changes no output)."""

  def __init__(self, batch_size: int, height: int, device, dtype=torch.uint8, capacity: int = 4):
    self.n, self.h, self.w = batch_size, height, 2 * height
    self.device = torch.device(device)
    self.count = 0
    self.masked = 0
    self._alloc(max(1, capacity), dtype)

  def _alloc(self, capacity, dtype):
    old = (self.rgb, self.depth, self.position) if self.count else None
    self.rgb = torch.zeros((self.n, capacity, self.h, self.w, 3), dtype=dtype, device=self.device)
    self.depth = torch.zeros((self.n, capacity, self.h, self.w), dtype=torch.float32, device=self.device)
    self.position = torch.zeros((self.n, capacity, 3), dtype=torch.float32, device=self.device)
    if old is not None:
      c = self.count
      self.rgb[:, :c] = old[0][:, :c].to(dtype)
      self.depth[:, :c] = old[1][:, :c]
      self.position[:, :c] = old[2][:, :c]

  @property
  def capacity(self):
    return self.rgb.shape[1]

  def _slot(self, masked: bool, dtype) -> int:
    """Makes room for one more frame and returns the slot it goes to."""
    want = torch.int32 if torch.int32 in (dtype, self.rgb.dtype) else torch.uint8
    if self.count == self.capacity or want != self.rgb.dtype:
      self._alloc(self.capacity * 2 if self.count == self.capacity else self.capacity, want)
    slot = self.count
    if masked and self.masked < self.count:  # keep the masked frames in front: move the first unmasked frame to the end
      m = self.masked
      self.rgb[:, slot] = self.rgb[:, m]
      self.depth[:, slot] = self.depth[:, m]
      self.position[:, slot] = self.position[:, m]
      slot = m
    self.count += 1
    self.masked += int(masked)
    return slot

  def append(self, rgb: torch.Tensor, depth: torch.Tensor, position: torch.Tensor, masked: bool = False) -> int:
    """rgb (N,H,W,3) uint8 / int32, depth (N,H,W[,1]) float32, position (N,3)."""
    slot = self._slot(masked, rgb.dtype)
    self.rgb[:, slot] = rgb.reshape(self.n, self.h, self.w, 3).to(self.rgb.dtype)
    self.depth[:, slot] = depth.reshape(self.n, self.h, self.w)
    self.position[:, slot] = position.reshape(self.n, 3)
    return slot

  def append_image(self, image: torch.Tensor, depth: torch.Tensor, position: torch.Tensor, masked: bool = False) -> int:
    """image (N,H,W,3) float32 in [0,1] (a ground-truth or generated frame): stored as
    clip(int32(image * 255), -1, 255) like the rollout loops do (trainers/gan_manager.py:539-542)."""
    slot = self._slot(masked, torch.int32)
    image = _lib.require_cuda(image, 'image').to(torch.float32)
    per_item = self.h * self.w * 3
    _lib.check(_lib.load().se3ds_quantize_rgb(_lib.ptr(image), self.n, per_item, _lib.ptr(self.rgb[:, slot]),
                                              self.capacity * per_item, _lib.stream_handle(self.device)))
    self.depth[:, slot] = depth.reshape(self.n, self.h, self.w)
    self.position[:, slot] = position.reshape(self.n, 3)
    return slot

  def reproject(self, tgt_pos, **kwargs) -> Dict[str, torch.Tensor]:
    if self.count == 0:
      raise ValueError('the frame ring is empty')
    return reproject(self.rgb, self.depth, self.position, tgt_pos, mask_frames=self.masked, frames=self.count, **kwargs)


def empty_memory_guidance(n: int, height: int, device) -> Dict[str, torch.Tensor]:
  """What the projection of an EMPTY memory yields (frame 0 of the rollout loops: tf.zeros((N,4,0))
  clouds, trainers/gan_manager.py:462-463): depth = clip(depth_scale) / depth_scale = 1, features = the
  output void class 0, mask 0."""
  w = 2 * height
  return dict(proj_image=torch.zeros((n, height, w, 3), device=device), proj_depth=torch.ones((n, height, w, 1), device=device),
              proj_mask=torch.zeros((n, height, w, 1), device=device))


def rollout(images, depths, positions, generator_fn: Callable, convention: Conventions = EVAL_METRIC,
            seq_len: Optional[int] = None, feed_depth: bool = True, depth_scale: float = constants.DEPTH_SCALE,
            workspace: Optional[_lib.Workspace] = None) -> Dict[str, list]:
  """Drop-in for the trajectory roll-out loops of the reference: trainers/gan_manager.py:458-556
  (`_get_image_grid`, mode "eval"; convention GAN_MANAGER) and utils/eval_metric.py:144-240
  (`_get_generated_pool`; convention EVAL_METRIC).  Per frame t:

    guidance(t)  = projection of the memory (frames 0 .. t-1) to positions[:, t]   -- empty at t = 0
    generated, depth_out = generator_fn(generator_inputs, t)                        -- the caller's generator
    memory      += frame 0: the ground-truth image with its blurred rows masked (-1), ground-truth depth
                   frame t > 0: the GENERATED image, and depth_out if feed_depth and it is not None

  images (N,T,H,W,3) float32 in [0,1], depths (N,T,H,W,1), positions (N,T,3).  generator_inputs holds
  prev_image, proj_image, proj_mask, proj_depth, blurred_mask, first_frame, exactly as the reference
  builds them; everything else the generator needs (one_hot_mask, dataset_type, ...) is the caller's
  closure.  generator_fn returns (generated (N,H,W,3) float32, depth_out (N,H,W,1) or None); tensors
  travel as torch CUDA tensors (DLPack-compatible, see INTEGRATION.md for the TF hand-off).
  Returns per-frame lists: guidance (dicts), generated, depth (the depth that went into the memory)."""
  images = _lib.require_cuda(torch.as_tensor(images), 'images').to(torch.float32)
  depths = _lib.require_cuda(torch.as_tensor(depths), 'depths').to(torch.float32)
  positions = _lib.require_cuda(torch.as_tensor(positions), 'positions').to(torch.float32)
  if images.dim() != 5 or images.shape[-1] != 3:
    raise ValueError(f'images should have shape (N, T, H, W, 3), got {tuple(images.shape)} instead.')
  n, t_all, h, w, _ = images.shape
  assert w == 2 * h, 'Expected equirectangular input images'
  steps = t_all if seq_len is None else int(seq_len)
  if not 0 < steps <= t_all:
    raise ValueError(f'seq_len must be in [1, {t_all}], got {seq_len}')
  depths = depths.reshape(n, t_all, h, w, 1)
  positions = positions.reshape(n, t_all, 3)
  dev = images.device
  ring = FrameRing(n, h, dev, torch.int32, capacity=steps)
  prev = torch.zeros_like(images[:, 0])
  result = dict(guidance=[], generated=[], depth=[])
  for t in range(steps):
    pos = positions[:, t]
    if ring.count == 0:
      g = empty_memory_guidance(n, h, dev)
    else:
      g = ring.reproject(pos, depth_scale=depth_scale, workspace=workspace, **convention.kwargs())
      g = {k: g[k] for k in ('proj_image', 'proj_depth', 'proj_mask')}
    g['blurred_mask'] = torch.zeros_like(g['proj_depth'])
    first = torch.ones((n,), device=dev) if t == 0 else torch.zeros((n,), device=dev)
    inputs = dict(prev_image=prev, proj_image=g['proj_image'], proj_mask=g['proj_mask'], proj_depth=g['proj_depth'],
                  blurred_mask=g['blurred_mask'], first_frame=first)
    generated, depth_out = generator_fn(inputs, t)
    generated = _lib.require_cuda(torch.as_tensor(generated), 'generated').to(torch.float32)
    depth_t = depths[:, t]
    if t == 0:  # ground truth, blurred rows masked (the ring masks its first `masked` frames in the kernel)
      prev = images[:, 0]
      ring.append_image(images[:, 0].contiguous(), depth_t, pos, masked=True)
    else:       # feed the generated frame back
      prev = generated
      if feed_depth and depth_out is not None:
        depth_t = _lib.require_cuda(torch.as_tensor(depth_out), 'depth_out').to(torch.float32).reshape(n, h, w, 1)
      ring.append_image(generated.contiguous(), depth_t, pos)
    result['guidance'].append(g)
    result['generated'].append(generated)
    result['depth'].append(depth_t)
  result['memory'] = ring
  return result


class MemoryState(NamedTuple):
  """The reference's memory tuple (models/models.py:77-87): coords (N,4,M) float32 + feats (N,M,1)
  uint8 = the semantic cloud, rgb_coords (N,4,M') float32 + rgb (N,M',3) int32 = the RGB cloud."""
  coords: torch.Tensor
  feats: torch.Tensor
  rgb_coords: torch.Tensor
  rgb: torch.Tensor


class FrameState(NamedTuple):
  """Native state of GuidanceMemory: the observed frames (see FrameRing)."""
  rgb: torch.Tensor         # (N,S,H,W,3) uint8 or int32
  semantic: torch.Tensor    # (N,S,H,W,1) uint8
  depth: torch.Tensor       # (N,S,H,W) float32
  position: torch.Tensor    # (N,S,3) float32
  masked: int               # the first `masked` frames get mask_pano(., 0.125, -1)


class GuidanceMemory(object):
  """The guidance half of the reference's SE3DSModel (models/models.py:90-321).

  add_to_memory(...) stores an observation; __call__(position) returns the generator inputs
  `proj_image`, `proj_depth`, `proj_mask`, `blurred_mask` (+ `proj_semantic`) for a target position --
  or for P positions at once (the VLN pose sweep, inference/perturbation_utils.py + notebook cell 13):
  position (P,3) gives P guidance sets from one call.  The generator itself
  (models/models.py:323-366) is out of scope.

  Natively the memory is a FrameRing.  `to_reference_state()` materialises the reference's
  MemoryState; `from_reference_state()` accepts one (e.g. a state saved by reference code): such a
  memory is a bare point cloud, so from then on the object works on clouds like the reference does
  (unproject + concatenate in add_to_memory, project_feats_to_equirectangular in __call__).
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
    self._ring: Optional[FrameRing] = None
    self._semantic: Optional[torch.Tensor] = None  # (N,cap,H,W,1) uint8, slots as in the ring
    self._cloud: Optional[MemoryState] = None      # cloud mode (after from_reference_state)

  # -- native state ---------------------------------------------------------------------------
  def get_memory_state(self) -> FrameState:
    if self._cloud is not None:
      raise ValueError('the memory holds a reference point cloud: use to_reference_state()')
    r = self._ring
    if r is None:
      raise ValueError('memory is empty: call add_to_memory first')
    c = r.count
    return FrameState(r.rgb[:, :c].clone(), self._semantic[:, :c].clone(), r.depth[:, :c].clone(),
                      r.position[:, :c].clone(), r.masked)

  def set_memory_state(self, state: FrameState):
    self.reset_memory()
    n, s = state.rgb.shape[:2]
    self._check_batch_size(n)
    r = self._ring = FrameRing(n, self.height, state.rgb.device, state.rgb.dtype, capacity=s)
    r.rgb.copy_(state.rgb); r.depth.copy_(state.depth); r.position.copy_(state.position)
    r.count, r.masked = s, int(state.masked)
    self._semantic = state.semantic.clone()

  # -- reference state (models/models.py:77-87,136-152) -----------------------------------------
  def _frame_cloud(self, feats, depth, pos, void):
    xyz1, f = pano_utils.equirectangular_to_pointcloud(feats, depth, void, self.depth_scale)
    return xyz1 + torch.cat([pos, torch.zeros_like(pos[:, :1])], dim=1)[:, :, None], f

  def to_reference_state(self) -> MemoryState:
    """The memory as the reference holds it: masking, unprojection, offset and compaction of every
    frame, concatenated (models/models.py:211-245).  Points appear frame by frame in ring order."""
    if self._cloud is not None:
      return MemoryState(*[t.clone() for t in self._cloud])
    r = self._ring
    dev = torch.device('cuda', torch.cuda.current_device()) if r is None else r.device
    coords, feats, rgb_coords, rgbs = [], [], [], []
    for i in range(0 if r is None else r.count):
      rgb = r.rgb[:, i].to(torch.int32)
      if i < r.masked:
        rgb = pano_utils.mask_pano(rgb, masked_region_value=constants.INVALID_RGB_VALUE)
      pos = r.position[:, i]
      xyz1, f = self._frame_cloud(rgb, r.depth[:, i], pos, constants.INVALID_RGB_VALUE)
      valid = (f != constants.INVALID_RGB_VALUE).any(dim=0).any(dim=-1)
      rgb_coords.append(xyz1[:, :, valid]); rgbs.append(f[:, valid])
      sxyz1, sf = self._frame_cloud(self._semantic[:, i], r.depth[:, i], pos, constants.INVALID_SEM_VALUE)
      svalid = (sf != constants.INVALID_SEM_VALUE).any(dim=0).any(dim=-1)
      coords.append(sxyz1[:, :, svalid]); feats.append(sf[:, svalid])
    n = self.batch_size
    if not coords:
      return MemoryState(torch.zeros((n, 4, 0), device=dev), torch.zeros((n, 0, 1), dtype=torch.uint8, device=dev),
                         torch.zeros((n, 4, 0), device=dev), torch.zeros((n, 0, 3), dtype=torch.int32, device=dev))
    return MemoryState(torch.cat(coords, dim=2), torch.cat(feats, dim=1), torch.cat(rgb_coords, dim=2), torch.cat(rgbs, dim=1))

  def from_reference_state(self, state: MemoryState):
    """Adopts a reference MemoryState (coords, feats, rgb_coords, rgb).  A cloud cannot be turned back
    into frames, so the memory switches to cloud mode (see the class docstring)."""
    self._check_batch_size(state.coords.shape[0])
    self.reset_memory()
    self._cloud = MemoryState(_lib.require_cuda(torch.as_tensor(state.coords), 'coords').to(torch.float32),
                              _lib.require_cuda(torch.as_tensor(state.feats), 'feats').to(torch.uint8),
                              _lib.require_cuda(torch.as_tensor(state.rgb_coords), 'rgb_coords').to(torch.float32),
                              _lib.require_cuda(torch.as_tensor(state.rgb), 'rgb').to(torch.int32))

  # -- the SE3DSModel interface -------------------------------------------------------------------
  def add_to_memory(self, pano_rgb, pano_semantic, pano_depth, position, mask_blurred=True):
    """models/models.py:180-245 (the cloud is not built; the frame is kept)."""
    pano_semantic = _lib.require_cuda(torch.as_tensor(pano_semantic), 'pano_semantic')
    self._check_batch_size(pano_semantic.shape[0])
    pano_rgb = _lib.require_cuda(torch.as_tensor(pano_rgb), 'pano_rgb')
    assert pano_rgb.dtype in (torch.uint8, torch.int32)
    assert pano_semantic.dtype in (torch.uint8, torch.int32)
    n = self.batch_size
    pano_depth = _lib.require_cuda(torch.as_tensor(pano_depth), 'pano_depth').to(torch.float32).reshape(n, self.height, self.width)
    position = _lib.require_cuda(torch.as_tensor(position), 'position').to(torch.float32).reshape(-1, 3)
    sem = pano_semantic.to(torch.uint8).reshape(n, self.height, self.width, 1)
    if self._cloud is not None:  # cloud mode: what the reference does
      rgb = pano_rgb.to(torch.int32)
      if mask_blurred:
        rgb = pano_utils.mask_pano(rgb, masked_region_value=constants.INVALID_RGB_VALUE)
      xyz1, f = self._frame_cloud(rgb, pano_depth, position, constants.INVALID_RGB_VALUE)
      valid = (f != constants.INVALID_RGB_VALUE).any(dim=0).any(dim=-1)
      sxyz1, sf = self._frame_cloud(sem, pano_depth, position, constants.INVALID_SEM_VALUE)
      svalid = (sf != constants.INVALID_SEM_VALUE).any(dim=0).any(dim=-1)
      c = self._cloud
      self._cloud = MemoryState(torch.cat([c.coords, sxyz1[:, :, svalid]], dim=2), torch.cat([c.feats, sf[:, svalid]], dim=1),
                                torch.cat([c.rgb_coords, xyz1[:, :, valid]], dim=2), torch.cat([c.rgb, f[:, valid]], dim=1))
      return
    if self._ring is None:
      self._ring = FrameRing(n, self.height, pano_rgb.device, pano_rgb.dtype)
      self._semantic = torch.zeros((n, self._ring.capacity, self.height, self.width, 1), dtype=torch.uint8, device=pano_rgb.device)
    r = self._ring
    count = r.count
    slot = r.append(pano_rgb, pano_depth, position, masked=bool(mask_blurred))
    if self._semantic.shape[1] != r.capacity:
      grown = torch.zeros((n, r.capacity, self.height, self.width, 1), dtype=torch.uint8, device=r.device)
      grown[:, :self._semantic.shape[1]] = self._semantic
      self._semantic = grown
    if slot != count:  # the ring moved the frame of `slot` to the end to keep the masked frames in front
      self._semantic[:, count] = self._semantic[:, slot]
    self._semantic[:, slot] = sem

  def __call__(self, position) -> Dict[str, torch.Tensor]:
    """The guidance half of models/models.py:247-321 for one target position (N,3) -- or for P target
    positions (P,3) / (N,P,3) at once: the leading dimension of every output is then N * P."""
    position = _lib.require_cuda(torch.as_tensor(position), 'position').to(torch.float32)
    position = position.reshape(self.batch_size, -1, 3)
    if self._cloud is not None:
      return self._call_cloud(position)
    r = self._ring
    if r is None or r.count == 0:
      raise ValueError('memory is empty: call add_to_memory first')
    out = r.reproject(position, depth_scale=self.depth_scale, unproject_void=SE3DS_MODEL.unproject_void,
                      project_void=SE3DS_MODEL.project_void, filter_void=True, per_job_bin=True)
    out['blurred_mask'] = torch.zeros_like(out['proj_mask'])
    if self.project_semantic:
      out['proj_semantic'] = self._project_semantic(position)
    return out

  def _call_cloud(self, position):
    """models/models.py:270-321 on the stored cloud, one reference call per target position."""
    c = self._cloud
    outs = []
    for p in range(position.shape[1]):
      pos = position[:, p]
      rel = c.rgb_coords - torch.cat([pos, torch.zeros_like(pos[:, :1])], dim=1)[:, :, None]
      d, f = pano_utils.project_feats_to_equirectangular(c.rgb, rel, self.height, self.width, constants.INVALID_RGB_VALUE,
                                                         self.depth_scale)
      mask = ((d > 0) & (d < 1) & (f != constants.INVALID_RGB_VALUE).all(dim=-1)).to(torch.float32)[..., None]
      # tensor / tensor is an IEEE division; tensor / python-scalar would be a multiplication by 1 / 255
      o = dict(proj_image=torch.clamp(f / torch.full((), 255.0, device=f.device), 0, 1), proj_depth=d[..., None], proj_mask=mask)
      if self.project_semantic:
        srel = c.coords - torch.cat([pos, torch.zeros_like(pos[:, :1])], dim=1)[:, :, None]
        _, sf = pano_utils.project_feats_to_equirectangular(c.feats, srel, self.height, self.width,
                                                            constants.INVALID_SEM_VALUE, self.depth_scale)
        o['proj_semantic'] = sf[..., 0].to(torch.uint8)
      outs.append(o)
    out = {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}
    out['blurred_mask'] = torch.zeros_like(out['proj_mask'])
    return out

  def rgb_cloud(self):
    """rgb_coords (N,4,M'), rgb (N,M',3) int32 of `to_reference_state()`."""
    st = self.to_reference_state()
    return st.rgb_coords, st.rgb

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
    r = self._ring
    sem3 = self._semantic[..., 0][..., None].expand(-1, -1, -1, -1, 3).contiguous()  # (N,cap,H,W,3)
    out = reproject(sem3, r.depth, r.position, position, self.depth_scale, mask_frames=0, frames=r.count,
                    unproject_void=constants.INVALID_SEM_VALUE, project_void=constants.INVALID_SEM_VALUE,
                    filter_void=True, per_job_bin=True, raw_features=True)
    return out['proj_image'][..., 0].to(torch.uint8)
