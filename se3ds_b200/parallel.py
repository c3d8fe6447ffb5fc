"""Multi-GPU form of the guidance path: one process per GPU, jobs sharded, one all-gather.

Every (batch item, target pose) job is independent, so the flattened job list is block
partitioned over the ranks and each rank runs the fused kernels on its shard with no data-path
collective.  The only exchange is the all-gather of the finished guidance
tensors -- the reference gathers per-replica outputs on the host
(trainers/gan_manager.py:612-615, utils/eval_metric.py:127-130) -- plus, for whole-call parity
of the global reject bin (utils/point_cloud_utils.py:150-153), a 4-float all-reduce.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import constants

_KEYS = ('proj_image', 'proj_depth', 'proj_mask', 'winner')


def shard_bounds(num_jobs: int, rank: int, world_size: int) -> Tuple[int, int]:
  """Contiguous block partition [lo, hi) of num_jobs; sizes differ by at most one."""
  base, rem = divmod(num_jobs, world_size)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def job_segments(lo: int, hi: int, num_poses: int) -> List[Tuple[int, int, int]]:
  """Splits the job range [lo, hi) (jobs are (n, p) row-major) into (n, p0, p1) segments."""
  segs = []
  j = lo
  while j < hi:
    n, p0 = divmod(j, num_poses)
    p1 = min(num_poses, p0 + (hi - j))
    segs.append((n, p0, p1))
    j += p1 - p0
  return segs


def _merge_whole_items(segs, num_poses):
  """Groups consecutive whole items so that one kernel call covers them."""
  groups, run = [], []
  for seg in segs:
    whole = seg[1] == 0 and seg[2] == num_poses
    if whole and run and run[-1][0] + 1 == seg[0]:
      run.append(seg)
    else:
      if run:
        groups.append(run)
      run = [seg] if whole else []
      if not whole:
        groups.append([seg])
  if run:
    groups.append(run)
  return groups


def reproject_sharded(rgb, depth, src_pos, tgt_pos, *, group=None, bin_mode: str = 'call',
                      gather: bool = True, compute_fn: Optional[Callable] = None,
                      apply_bin_fn: Optional[Callable] = None, depth_scale: float = constants.DEPTH_SCALE,
                      **kwargs) -> Dict[str, torch.Tensor]:
  """Shards N*P jobs over the ranks of `group`, runs the fused path, all-gathers the guidance.

  Every rank passes the same full inputs (or at least its own shard's items; other items are
  never read).  bin_mode: 'call' = the reference's whole-call reject bin (pixel (0,0) of global
  job 0, exact: a 4-float all-reduce), 'shard' = each rank's own first job (what the reference
  does per replica under MirroredStrategy, trainers/gan_manager.py:577), 'job' = per job.
  Returns the dict of `guidance.reproject`; with gather=True the tensors cover all J jobs on
  every rank, otherwise only the local shard ('job_range' tells which).
  compute_fn / apply_bin_fn exist so that the host logic can be exercised without a GPU.
  """
  if compute_fn is None:
    from . import guidance
    compute_fn = guidance.reproject
    apply_bin_fn = guidance.apply_bin
  rank = dist.get_rank(group) if dist.is_initialized() else 0
  world = dist.get_world_size(group) if dist.is_initialized() else 1
  rgb = torch.as_tensor(rgb)
  if rgb.dim() == 4:
    rgb = rgb[:, None]
  n = rgb.shape[0]
  tgt_pos = torch.as_tensor(tgt_pos).reshape(n, -1, 3)
  p = tgt_pos.shape[1]
  lo, hi = shard_bounds(n * p, rank, world)
  depth = torch.as_tensor(depth).reshape(rgb.shape[:4])
  src_pos = torch.as_tensor(src_pos).reshape(n, rgb.shape[1], 3)

  outs, bins = [], []
  for grp in _merge_whole_items(job_segments(lo, hi, p), p):
    n0, n1 = grp[0][0], grp[-1][0] + 1
    p0, p1 = grp[0][1], grp[0][2]
    o = compute_fn(rgb[n0:n1], depth[n0:n1], src_pos[n0:n1], tgt_pos[n0:n1, p0:p1], depth_scale=depth_scale,
                   per_job_bin=(bin_mode == 'job'), export_bin=(bin_mode in ('call', 'shard')), **kwargs)
    if bin_mode in ('call', 'shard'):
      bins.append(o.pop('bin'))
    outs.append(o)
  h, w = rgb.shape[2], rgb.shape[3]
  if outs:
    local = {k: torch.cat([o[k] for o in outs], dim=0) for k in _KEYS if k in outs[0]}
  else:  # a rank without jobs still takes part in the collectives
    dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
    local = dict(proj_image=torch.empty((0, h, w, 3), device=dev), proj_depth=torch.empty((0, h, w, 1), device=dev),
                 proj_mask=torch.empty((0, h, w, 1), device=dev))
    if kwargs.get('return_winner'):
      local['winner'] = torch.empty((0, h, w), dtype=torch.int32, device=dev)
  dev = local['proj_image'].device

  # reject bin (min depth, max R, G, B): one MAX all-reduce of (-zmin, r, g, b)
  if bin_mode in ('call', 'shard'):
    if bins:
      b = torch.stack(bins)
      red = torch.cat([(-b[:, :1]).max(dim=0).values, b[:, 1:4].max(dim=0).values])
    else:
      red = torch.tensor([-float('inf'), 0.0, 0.0, 0.0], device=dev)
    if bin_mode == 'call' and world > 1:
      dist.all_reduce(red, op=dist.ReduceOp.MAX, group=group)
    owner = (lo == 0 and hi > 0) if bin_mode == 'call' else hi > lo
    if owner:
      # fifth value: depth of the owner pixel's own winner, from the call that rendered this rank's first job
      red = torch.cat([-red[:1], red[1:], bins[0][4:5].to(red.device)])
      if kwargs.get('raw_features'):
        apply_bin_fn(red, local, depth_scale, raw_features=True)
      else:
        apply_bin_fn(red, local, depth_scale)

  result = dict(local)
  result['job_range'] = (lo, hi)
  if gather and world > 1:
    # equal-sized (padded) shards -> a single all_gather_into_tensor per guidance tensor
    cap = -(-(n * p) // world)
    for k, v in local.items():
      send = v.contiguous()
      if send.shape[0] < cap:
        send = torch.cat([send, send.new_zeros((cap - send.shape[0],) + tuple(send.shape[1:]))], dim=0)
      recv = send.new_empty((world * cap,) + tuple(send.shape[1:]))
      dist.all_gather_into_tensor(recv, send, group=group)
      parts = []
      for r in range(world):
        a, b = shard_bounds(n * p, r, world)
        parts.append(recv[r * cap:r * cap + (b - a)])
      result[k] = torch.cat(parts, dim=0)
  return result
