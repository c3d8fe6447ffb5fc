"""ctypes binding of libse3ds_geom.so (include/se3ds_geom.h).

torch is used only for device memory and streams: every call below passes raw
device pointers (`tensor.data_ptr()`) and the current CUDA stream handle to the
C ABI.  There is no CPU fallback: if the library cannot be loaded or no CUDA
device is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Dict, Optional

import torch

from . import build as _build

OK, ERR_BAD_SHAPE, ERR_BAD_DTYPE, ERR_BAD_ARG, ERR_CUDA, ERR_NOMEM = range(6)
U8, I32, F32 = 0, 1, 2
FLAG_FILTER_VOID = 1
FLAG_BIN_PER_JOB = 2
FLAG_KEY64 = 4
FLAG_INPUTS_READY = 8
FLAG_RAW_FEATURES = 16
FLAG_COMPACT_OUT = 32
FLAG_HOST_ASYNC = 64

_DTYPES = {torch.uint8: U8, torch.int32: I32, torch.float32: F32}

_lib = None
_lock = threading.Lock()

_c = ctypes
_vp, _i, _ll, _f, _d, _u, _sz = _c.c_void_p, _c.c_int, _c.c_longlong, _c.c_float, _c.c_double, _c.c_uint, _c.c_size_t

# name -> argtypes; must list every symbol include/se3ds_geom.h declares
SIGNATURES = {
    'se3ds_version': [],
    'se3ds_status_string': [_i],
    'se3ds_last_error': [],
    'se3ds_ws_create': [_i, _sz, _sz, _c.POINTER(_vp)],
    'se3ds_ws_destroy': [_vp],
    'se3ds_ws_bytes': [_vp, _c.POINTER(_sz)],
    'se3ds_ws_projection_mode': [_vp, _i, _f],
    'se3ds_ws_pdl': [_vp, _i],
    'se3ds_ws_lanes': [_vp, _i, _ll, _i],
    'se3ds_plan_chunks': [_sz, _i, _ll, _i, _i, _i, _i, _i, _i, _c.POINTER(_ll * 5)],
    'se3ds_ws_verify_read': [_vp, _c.POINTER(_c.c_ulonglong * 3), _c.POINTER(_f * 2)],
    'se3ds_ws_profile': [_vp, _i],
    'se3ds_ws_profile_read': [_vp, _c.POINTER(_f * 3), _c.POINTER(_c.c_ulonglong)],
    'se3ds_ws_profile_read_stamps': [_vp, _c.POINTER(_d * 3), _c.POINTER(_ll)],
    'se3ds_mask_pano': [_vp, _i, _i, _i, _i, _i, _d, _d, _vp, _vp],
    'se3ds_unproject_equirect': [_vp, _vp, _i, _vp, _i, _i, _i, _i, _d, _f, _vp, _vp, _i, _vp],
    'se3ds_project_cloud': [_vp, _vp, _vp, _i, _i, _ll, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp],
    'se3ds_reproject': [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _d, _i, _i, _i, _u, _vp, _vp, _vp,
                        _vp, _vp, _vp],
    'se3ds_reproject_se3': [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _d, _i, _i, _i, _u, _vp, _vp,
                            _vp, _vp, _vp, _vp],
    'se3ds_reproject_ring': [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _d, _i, _i, _i, _u, _vp,
                             _vp, _vp, _vp, _vp, _vp],
    'se3ds_expand_guidance': [_vp, _vp, _ll, _ll, _vp, _vp, _vp, _vp, _vp],
    'se3ds_quantize_rgb': [_vp, _i, _ll, _vp, _ll, _vp],
    'se3ds_apply_bin': [_vp, _f, _u, _vp, _vp, _vp, _vp, _vp],
    'se3ds_reproject_host': [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _d, _i, _i, _i, _u, _vp, _vp,
                             _vp, _vp],
    'se3ds_ws_host_wait': [_vp],
    'se3ds_resize': [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    'se3ds_interpolate_bilinear': [_vp, _vp, _i, _i, _i, _i, _ll, _i, _vp, _vp],
    'se3ds_filtered_coords_and_feats': [_vp, _i, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp],
    'se3ds_pixel_rays': [_i, _vp, _vp],
    'se3ds_rotate_pano': [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    'se3ds_project_perspective_image': [_vp, _i, _i, _i, _c.POINTER(_f * 9), _i, _i, _f, _i, _vp, _vp],
    'se3ds_perspective_from_equirect': [_vp, _i, _i, _i, _c.POINTER(_f * 9), _c.POINTER(_f * 9), _i, _i, _vp, _vp],
    'se3ds_proportion_invalid': [_vp, _i, _vp, _i, _i, _f, _f, _vp, _vp],
}


def library_path() -> str:
  return _build.OUT


def load():
  """Loads (building first if the .so is missing and nvcc is available) the C-ABI library."""
  global _lib
  with _lock:
    if _lib is not None:
      return _lib
    path = library_path()
    if not os.path.exists(path):
      try:
        _build.build()
      except Exception as e:  # pylint: disable=broad-except
        raise RuntimeError(
            f'{path} is missing and could not be built ({e}); run `python -m se3ds_b200.build`. '
            'There is no CPU fallback for the geometric guidance path.') from e
    lib = ctypes.CDLL(path)
    for name, argtypes in SIGNATURES.items():
      fn = getattr(lib, name)
      fn.argtypes = argtypes
      fn.restype = ctypes.c_char_p if name in ('se3ds_status_string', 'se3ds_last_error') else ctypes.c_int
    _lib = lib
    return lib


class Se3dsError(RuntimeError):
  pass


def check(status: int):
  if status == OK:
    return
  lib = load()
  msg = (lib.se3ds_last_error() or b'').decode() or lib.se3ds_status_string(status).decode()
  if status in (ERR_BAD_SHAPE, ERR_BAD_DTYPE, ERR_BAD_ARG):
    raise ValueError(msg)
  if status == ERR_NOMEM:
    raise torch.cuda.OutOfMemoryError(msg)
  raise Se3dsError(msg)


def dtype_code(t: torch.Tensor) -> int:
  try:
    return _DTYPES[t.dtype]
  except KeyError as e:
    raise ValueError(f'unsupported tensor dtype {t.dtype}') from e


def require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
  if not isinstance(t, torch.Tensor):
    raise TypeError(f'{name} must be a torch.Tensor, got {type(t)}')
  if not t.is_cuda:
    if not torch.cuda.is_available():
      raise Se3dsError('no CUDA device: the geometric guidance path has no CPU fallback')
    t = t.cuda()
  return t.contiguous()


def ptr(t: Optional[torch.Tensor]):
  return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_handle(device) -> ctypes.c_void_p:
  return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Workspace:
  """Opaque se3ds_ws handle: z-buffer, feature buffer, scratch, bins, angle tables."""

  def __init__(self, device: int = 0, max_bytes: int = 0, l2_chunk_bytes: int = 0):
    self._h = ctypes.c_void_p()
    self.device = device
    check(load().se3ds_ws_create(device, max_bytes, l2_chunk_bytes, ctypes.byref(self._h)))

  @property
  def handle(self):
    if not self._h:
      raise Se3dsError('workspace was destroyed')
    return self._h

  def nbytes(self) -> int:
    out = ctypes.c_size_t()
    check(load().se3ds_ws_bytes(self.handle, ctypes.byref(out)))
    return out.value

  def projection_mode(self, mode: int, margin_scale: float = 0.0):
    """0 canonical only, 1 certified fast path (default), 2 verify (see se3ds_geom.h)."""
    check(load().se3ds_ws_projection_mode(self.handle, int(mode), float(margin_scale)))

  def pdl(self, enable: bool):
    """Programmatic dependent launch between the fused kernels (default on)."""
    check(load().se3ds_ws_pdl(self.handle, int(enable)))

  def lanes(self, lanes: int, min_points_per_lane: int = 0, min_chunks_per_lane: int = 0):
    """Concurrent chunk lanes of one reproject call (1..4, default 2; 0 = default lane minima)."""
    check(load().se3ds_ws_lanes(self.handle, int(lanes), int(min_points_per_lane), int(min_chunks_per_lane)))

  def verify_read(self):
    """-> dict(points, certified, wrong, max_dev_x, max_dev_y) accumulated in verify mode."""
    c = (ctypes.c_ulonglong * 3)()
    d = (ctypes.c_float * 2)()
    check(load().se3ds_ws_verify_read(self.handle, ctypes.byref(c), ctypes.byref(d)))
    return dict(points=c[0], certified=c[1], wrong=c[2], max_dev_x=d[0], max_dev_y=d[1])

  def host_wait(self):
    """Completes a pending asynchronous se3ds_reproject_host call of this workspace."""
    check(load().se3ds_ws_host_wait(self.handle))

  def profile(self, mode):
    """0 / False off, 1 / True cudaEvents between the launches (no overlap), 2 end-of-kernel stamps."""
    check(load().se3ds_ws_profile(self.handle, int(mode)))

  def profile_read_stamps(self):
    """-> ((ms K2, K3, K4) summed shares of the pipelined step, chunks counted) since profile(2)."""
    ms = (ctypes.c_double * 3)()
    n = ctypes.c_longlong()
    check(load().se3ds_ws_profile_read_stamps(self.handle, ctypes.byref(ms), ctypes.byref(n)))
    return tuple(ms), n.value

  def profile_read(self):
    """-> ((ms_splat_depth, ms_splat_feat, ms_resolve) since the last read, kernels launched so far)."""
    ms = (ctypes.c_float * 3)()
    n = ctypes.c_ulonglong()
    check(load().se3ds_ws_profile_read(self.handle, ctypes.byref(ms), ctypes.byref(n)))
    return tuple(ms), n.value

  def close(self):
    if self._h:
      load().se3ds_ws_destroy(self._h)
      self._h = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


_default_ws: Dict[tuple, Workspace] = {}


def plan_chunks(n: int, s: int, p: int, h: int, l2_chunk_bytes: int = 0, lanes: int = 2, min_points_per_lane: int = 0,
                min_chunks_per_lane: int = 0) -> dict:
  """The chunk / lane plan of a reproject call of this shape (host arithmetic only, works without a GPU)."""
  out = (_ll * 5)()
  check(load().se3ds_plan_chunks(int(l2_chunk_bytes), int(lanes), int(min_points_per_lane), int(min_chunks_per_lane),
                                 int(n), int(s), int(p), int(h), int(2 * h), ctypes.byref(out)))
  return dict(lanes=out[0], items_per_chunk=out[1], poses_per_chunk=out[2], chunk_jobs=out[3], nchunks=out[4])


def default_workspace(device, role: str = 'stream') -> Workspace:
  """The process-wide default workspace for `device` AND the CUDA stream that is current on it.

  The C ABI allows a workspace on one stream at a time (its z-buffer, feature buffer, scratch and bins
  are reused by every call), so calls issued on different torch streams -- or through
  se3ds_reproject_host, which runs on the workspace's own internal streams (role='host') -- must not
  share one: the default is keyed by (device, stream handle / role).  Pass an explicit Workspace to
  control placement yourself."""
  idx = torch.device(device).index
  if idx is None:
    idx = torch.cuda.current_device()
  key = (idx, 'host') if role == 'host' else (idx, int(torch.cuda.current_stream(idx).cuda_stream))
  ws = _default_ws.get(key)
  if ws is None:
    chunk = int(os.environ.get('SE3DS_L2_CHUNK_MB', '0')) << 20
    ws = _default_ws[key] = Workspace(idx, 0, chunk)
  return ws
