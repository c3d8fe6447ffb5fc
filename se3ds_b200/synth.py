"""Synthetic RGB-D panoramas for tests and benchmarks: numpy, seeded.

D_room: analytic distance from a camera inside an axis-aligned 8 x 6 x 3 m box to its walls,
  / 20, with 2 % multiplicative noise, 3 % of the pixels 0 and 1 % exactly 1.0 (both invalid).
D_rand: i.i.d. U[0,1) depth as in the reference tests (utils/pano_utils_test.py:99-102).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
ROOM = np.array([8.0, 6.0, 3.0])


def pixel_directions(h: int, w: int):
  """Unit ray of every equirect pixel in the reference's convention (pano_utils.py:211-236)."""
  hp = 0.5 * np.pi / h
  elev = np.linspace(hp, np.pi - hp, h)
  head = np.linspace(1.5 * np.pi - hp, -0.5 * np.pi + hp, w)
  se, ce = np.sin(elev)[:, None], np.cos(elev)[:, None]
  return np.stack([se * np.cos(head)[None, :], se * np.sin(head)[None, :], np.broadcast_to(ce, (h, w))], axis=-1)


def room_depth(h: int, w: int, cam: np.ndarray, rng: np.random.Generator, noise: bool = True) -> np.ndarray:
  d = pixel_directions(h, w)
  with np.errstate(divide='ignore', invalid='ignore'):
    t_hi = (ROOM[None, None, :] - cam) / d
    t_lo = (0.0 - cam) / d
  t = np.where(d > 0, t_hi, np.where(d < 0, t_lo, np.inf)).min(axis=-1)
  depth = t / 20.0
  if noise:
    depth = depth * (1 + 0.02 * rng.uniform(-1, 1, depth.shape))
    u = rng.uniform(0, 1, depth.shape)
    depth = np.where(u < 0.03, 0.0, np.where(u > 0.99, 1.0, depth))
  return depth.astype(F32)


def make_inputs(n: int, s: int, p: int, h: int, seed: int = 0, dist: str = 'room', rgb_dtype=np.uint8,
                sweep: bool = False):
  """Returns dict(rgb (N,S,H,W,3), depth (N,S,H,W), src_pos (N,S,3), tgt_pos (N,P,3)).

  Source positions follow a trajectory p_0 + k*(1.4, 0.1, 0); the target is the last source
  position + (1.0, 0.3, 0.05), or -- sweep=True -- P draws of U([-1.5,1.5]^2 x [-0.1,0.1])
  around it (notebooks/SE3DS_VLN_Augmentation_Colab.ipynb cell 13).
  """
  rng = np.random.default_rng(seed)
  w = 2 * h
  rgb = rng.integers(0, 256, (n, s, h, w, 3), dtype=np.uint8).astype(rgb_dtype)
  depth = np.empty((n, s, h, w), F32)
  src = np.empty((n, s, 3), F32)
  step = np.array([1.4, 0.1, 0.0])
  for b in range(n):
    # trajectory start so that every frame stays >= 0.5 m inside the room
    span = step * (s - 1)
    p0 = np.array([rng.uniform(0.5, max(0.5, ROOM[0] - 0.5 - span[0])),
                   rng.uniform(0.5, max(0.5, ROOM[1] - 0.5 - span[1])), rng.uniform(0.5, ROOM[2] - 0.5)])
    for k in range(s):
      cam = np.minimum(p0 + k * step, ROOM - 0.5)
      src[b, k] = cam
      depth[b, k] = room_depth(h, w, cam, rng) if dist == 'room' else rng.uniform(0, 1, (h, w)).astype(F32)
  if sweep:
    noise = np.concatenate([rng.uniform(-1.5, 1.5, (n, p, 2)), rng.uniform(-0.1, 0.1, (n, p, 1))], axis=-1)
    tgt = src[:, -1:, :] + noise
  else:
    tgt = np.repeat(src[:, -1:, :] + np.array([1.0, 0.3, 0.05]), p, axis=1)
    if p > 1:
      tgt = tgt + rng.uniform(-0.2, 0.2, tgt.shape)
  return dict(rgb=rgb, depth=depth, src_pos=src.astype(F32), tgt_pos=tgt.astype(F32))
