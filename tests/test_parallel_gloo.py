"""world_size-2 gloo test of the multi-GPU host logic (se3ds_b200/parallel.py) on CPU.

The CUDA kernels cannot run here, so the per-shard compute function is a numpy emulation of the
export-bin protocol built on the canonical oracle; what is under test is the sharding, the bin
all-reduce, the owner patch and the padded all-gather.  The gathered result must equal the
single-call oracle bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_exact as X
from se3ds_b200 import parallel, synth

F32 = np.float32


def oracle_compute(rgb, depth, src_pos, tgt_pos, depth_scale=20.0, per_job_bin=False, export_bin=False,
                   return_winner=False, out=None, compact=False, **kw):
  """Emulates guidance.reproject(export_bin=..., compact=...) with the oracle (CPU tensors); `out` is ignored
  (fresh tensors are returned, reproject_sharded copies them into its gather buffer)."""
  rgb, depth, src_pos, tgt_pos = (np.asarray(t) for t in (rgb, depth, src_pos, tgt_pos))
  o = X.reproject(rgb, depth, src_pos, tgt_pos, mask_first_frame=False, per_job_bin=per_job_bin)
  out = dict(proj_image=o['image'].copy(), proj_depth=o['depth'].copy(), proj_mask=o['mask'].copy())
  if return_winner:
    out['winner'] = o['winner']
  if export_bin:
    j, h, w = o['winner'].shape
    flat, rad, valid = o['flat'].reshape(-1), o['rad'].reshape(-1), o['valid'].reshape(-1).astype(bool)
    feats = o['feats_in'].reshape(-1, 3)
    at0 = valid & (flat == 0)
    z0 = min(F32(depth_scale), rad[at0].min()) if at0.any() else F32(depth_scale)
    zb = o['zbuf'].reshape(-1).copy()
    zb[0] = z0
    keep = valid & (rad < zb[flat] + F32(0.1))
    binpts = ~keep
    binz = rad[~valid].min() if (~valid).any() else np.inf
    binf = np.maximum(feats[binpts].max(axis=0), 0) if binpts.any() else np.zeros(3, F32)
    f0 = np.maximum(feats[keep & (flat == 0)].max(axis=0), 0) if (keep & (flat == 0)).any() else np.zeros(3, F32)
    d0 = np.clip(z0, 0, F32(depth_scale)) / F32(depth_scale)
    out['proj_depth'][0, 0, 0, 0] = d0
    out['proj_image'][0, 0, 0] = np.clip(f0 / F32(255), 0, 1)
    out['proj_mask'][0, 0, 0, 0] = F32(0 < d0 < 1)
    # the owner pixel's own winner ignores the (not yet reduced) reject bin; its depth travels with the bin
    own = np.inf
    if at0.any() and rad[at0].min() <= F32(depth_scale):
      own = rad[at0].min()
      if return_winner:
        m = o['flat'].reshape(j, -1).shape[1]
        out['winner'] = out['winner'].copy()
        out['winner'][0, 0, 0] = int(np.flatnonzero(at0 & (rad == own))[0] % m)
    elif return_winner:
      out['winner'] = out['winner'].copy()
      out['winner'][0, 0, 0] = -1
    out['bin'] = np.array([binz, *binf, own], F32)
  if compact:  # uint8 colours + depth; the mask is a function of the depth
    out['proj_rgb_u8'] = np.rint(out.pop('proj_image') * F32(255)).astype(np.uint8)
    out.pop('proj_mask')
  return {k: torch.as_tensor(v) for k, v in out.items()}


def oracle_expand(out, job_map=None):
  """Emulates guidance.expand_guidance."""
  rgb8, depth = out['proj_rgb_u8'].numpy(), out['proj_depth'].numpy()
  image = np.clip(rgb8.astype(F32) / F32(255), 0, 1).astype(F32)
  mask = ((depth > 0) & (depth < 1)).astype(F32)
  if job_map is not None:
    inv = np.empty_like(job_map.numpy()); inv[job_map.numpy()] = np.arange(len(inv))
    image, mask, depth = image[inv], mask[inv], depth[inv]
    out['proj_depth'] = torch.as_tensor(depth)
    out.pop('proj_rgb_u8')
    if 'winner' in out:
      out['winner'] = out['winner'][torch.as_tensor(inv)]
  out['proj_image'], out['proj_mask'] = torch.as_tensor(image), torch.as_tensor(mask)
  return out


def oracle_apply_bin(bin_values, out, depth_scale=20.0):
  b = bin_values.numpy().astype(F32)
  d = min(out['proj_depth'][0, 0, 0, 0].item(), float(np.clip(b[0], 0, F32(depth_scale)) / F32(depth_scale)))
  out['proj_depth'][0, 0, 0, 0] = d
  if 'proj_rgb_u8' in out and 'proj_image' not in out:
    img8 = np.maximum(out['proj_rgb_u8'][0, 0, 0].numpy(), np.clip(b[1:4], 0, 255).astype(np.uint8))
    out['proj_rgb_u8'][0, 0, 0] = torch.as_tensor(img8)
  else:
    img = np.maximum(out['proj_image'][0, 0, 0].numpy(), np.clip(b[1:4] / F32(255), 0, 1))
    out['proj_image'][0, 0, 0] = torch.as_tensor(img)
    out['proj_mask'][0, 0, 0, 0] = float(0 < d < 1)
  if 'winner' in out and b[0] < b[4]:
    out['winner'][0, 0, 0] = -1


def _worker(rank, world, port, n, p, bin_mode, wire, chunks, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    inp = synth.make_inputs(n, 2, p, 16, seed=3, dist='rand', sweep=p > 1)
    res = parallel.reproject_sharded(inp['rgb'], inp['depth'], inp['src_pos'], inp['tgt_pos'], bin_mode=bin_mode,
                                     wire=wire, chunks=chunks, compute_fn=oracle_compute, apply_bin_fn=oracle_apply_bin,
                                     expand_fn=oracle_expand, return_winner=True)
    q.put((rank, {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in res.items()}))
  finally:
    dist.destroy_process_group()


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


@pytest.mark.parametrize('n,p,bin_mode,wire,chunks', [
    (4, 1, 'call', 'compact', 0), (1, 5, 'call', 'compact', 0), (3, 3, 'call', 'f32', 0), (3, 2, 'job', 'compact', 0),
    (1, 1, 'call', 'compact', 0), (1, 8, 'call', 'compact', 2), (4, 2, 'call', 'compact', 4), (2, 2, 'call', 'f32', 1)])
def test_sharded_equals_single_call(n, p, bin_mode, wire, chunks):
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, n, p, bin_mode, wire, chunks, q)) for r in range(world)]
  for pr in procs:
    pr.start()
  results = dict(q.get(timeout=240) for _ in range(world))
  for pr in procs:
    pr.join(timeout=60)
    assert pr.exitcode == 0
  inp = synth.make_inputs(n, 2, p, 16, seed=3, dist='rand', sweep=p > 1)
  ref = X.reproject(inp['rgb'], inp['depth'], inp['src_pos'], inp['tgt_pos'], mask_first_frame=False,
                    per_job_bin=(bin_mode == 'job'))
  ranges = sorted(results[r]['job_range'] for r in range(world))
  assert ranges[0][0] == 0 and ranges[-1][1] == n * p and ranges[0][1] == ranges[1][0]
  for r in range(world):
    np.testing.assert_array_equal(results[r]['proj_image'], ref['image'])
    np.testing.assert_array_equal(results[r]['proj_depth'], ref['depth'])
    np.testing.assert_array_equal(results[r]['proj_mask'], ref['mask'])
    np.testing.assert_array_equal(results[r]['winner'], ref['winner'])


def test_shard_bounds_and_segments():
  for jobs in (0, 1, 7, 64):
    for world in (1, 2, 4, 8):
      b = [parallel.shard_bounds(jobs, r, world) for r in range(world)]
      assert b[0][0] == 0 and b[-1][1] == jobs
      assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
      assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
  assert parallel.job_segments(3, 9, 4) == [(0, 3, 4), (1, 0, 4), (2, 0, 1)]
  assert parallel._merge_whole_items(parallel.job_segments(0, 12, 4), 4) == [[(0, 0, 4), (1, 0, 4), (2, 0, 4)]]
