"""Multi-GPU form of the guidance path: one process per GPU, jobs sharded, one all-gather.

Every (batch item, target pose) job is independent, so the flattened job list is block
partitioned over the ranks and each rank runs the fused kernels on its shard with no data-path
collective.  The only exchange is the all-gather of the finished guidance
tensors -- the reference gathers per-replica outputs on the host
(trainers/gan_manager.py:612-615, utils/eval_metric.py:127-130) -- plus, for whole-call parity
of the global reject bin (utils/point_cloud_utils.py:150-153), a 5-float all-reduce.  The gather ships a
compact wire format (uint8 colours + float32 depth, 7 B per pixel instead of 20) and runs in place in a
pre-sized buffer, piece by piece while the next piece renders (see reproject_sharded).
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


def _chunk_groups(groups, chunks: int):
  """Deals the kernel-call groups of a shard into `chunks` consecutive runs (for the gather pipeline)."""
  chunks = max(1, min(chunks, len(groups)))
  base, rem = divmod(len(groups), chunks)
  out, i = [], 0
  for c in range(chunks):
    k = base + (1 if c < rem else 0)
    out.append(groups[i:i + k])
    i += k
  return out


def reproject_sharded(rgb, depth, src_pos, tgt_pos, *, group=None, bin_mode: str = 'call',
                      gather: bool = True, wire: str = 'compact', chunks: int = 0, expand: bool = True,
                      compute_fn: Optional[Callable] = None, apply_bin_fn: Optional[Callable] = None,
                      expand_fn: Optional[Callable] = None, depth_scale: float = constants.DEPTH_SCALE,
                      **kwargs) -> Dict[str, torch.Tensor]:
  """Shards N*P jobs over the ranks of `group`, runs the fused path, all-gathers the guidance.

  Every rank passes the same full inputs (or at least its own shard's items; other items are
  never read).  bin_mode: 'call' = the reference's whole-call reject bin (pixel (0,0) of global
  job 0, exact: a 5-float all-reduce), 'shard' = each rank's own first job (what the reference
  does per replica under MirroredStrategy, trainers/gan_manager.py:577), 'job' = per job.
  Returns the dict of `guidance.reproject`; with gather=True the tensors cover all J jobs on
  every rank, otherwise only the local shard ('job_range' tells which).

  The gather (the reference concatenates per-replica outputs on the host, trainers/gan_manager.py:612-615,
  utils/eval_metric.py:255-266): every rank renders straight into its slot of one pre-sized gather buffer
  per tensor and the all-gather runs in place -- no padding copy before, no concatenation after (only a
  job count that does not divide by the world size costs one compaction).  wire='compact' (default)
  ships uint8 colours + float32 depth (7 B per pixel instead of 20; the mask is a function of the depth)
  and expands to the float32 tensors on arrival, bit-identical by construction (`expand=False` keeps the
  compact tensors); wire='f32' ships the float32 tensors themselves.  With chunks > 1 the shard is rendered
  in that many pieces and the collective of piece k travels (async, NCCL's stream) while piece k+1
  renders; chunks=0 picks min(4, kernel calls of the shard).  The reject bin is reduced with one tiny
  all-reduce and applied after the gather, on every rank's copy.
  compute_fn / apply_bin_fn / expand_fn exist so that the host logic can be exercised without a GPU.
  """
  if compute_fn is None:
    from . import guidance
    compute_fn = guidance.reproject
    apply_bin_fn = guidance.apply_bin
    expand_fn = guidance.expand_guidance
  if wire not in ('compact', 'f32'):
    raise ValueError(f"wire must be 'compact' or 'f32', got {wire!r}")
  if bin_mode not in ('call', 'shard', 'job'):
    raise ValueError(f"bin_mode must be 'call', 'shard' or 'job', got {bin_mode!r}")
  raw = bool(kwargs.get('raw_features'))
  compact = wire == 'compact' and gather and not raw
  want_winner = bool(kwargs.get('return_winner'))
  rank = dist.get_rank(group) if dist.is_initialized() else 0
  world = dist.get_world_size(group) if dist.is_initialized() else 1
  rgb = torch.as_tensor(rgb)
  if rgb.dim() == 4:
    rgb = rgb[:, None]
  n = rgb.shape[0]
  tgt_pos = torch.as_tensor(tgt_pos).reshape(n, -1, 3)
  p = tgt_pos.shape[1]
  jobs = n * p
  lo, hi = shard_bounds(jobs, rank, world)
  depth = torch.as_tensor(depth).reshape(rgb.shape[:4])
  src_pos = torch.as_tensor(src_pos).reshape(n, rgb.shape[1], 3)
  h, w = rgb.shape[2], rgb.shape[3]
  on_gpu = torch.cuda.is_available() and compute_fn.__module__.startswith('se3ds_b200')
  dev = torch.device('cuda', torch.cuda.current_device()) if on_gpu else torch.device('cpu')

  # One gather buffer per tensor.  Layout: pieces x world x (jobs per piece): piece c of every rank is one
  # contiguous block, so its all-gather runs in place while piece c + 1 renders.  With one piece that is
  # rank-major, i.e. job order (up to the padding of ranks that hold one job less).
  multi = gather and world > 1
  cap = -(-jobs // world) if multi else hi - lo
  npieces = 1
  if multi and compact and expand and jobs % world == 0 and cap > 1:
    npieces = min(cap, chunks if chunks > 0 else 4)
    while cap % npieces:
      npieces -= 1
  jpp = cap // npieces                      # jobs per piece
  slots = npieces * world * jpp if multi else cap
  keys = {'proj_rgb_u8': ((h, w, 3), torch.uint8), 'proj_depth': ((h, w, 1), torch.float32)} if compact else \
         {'proj_image': ((h, w, 3), torch.float32), 'proj_depth': ((h, w, 1), torch.float32), 'proj_mask': ((h, w, 1), torch.float32)}
  if want_winner:
    keys['winner'] = ((h, w), torch.int32)
  bufs = {k: torch.empty((slots,) + shp, dtype=dt, device=dev) for k, (shp, dt) in keys.items()}

  def slot_of(local_job):                   # slot of this rank's local job in the gather buffer
    c, j = divmod(local_job, jpp)
    return ((c * world + rank) * jpp + j) if multi else local_job

  bins, handles = [], []
  for c in range(npieces):
    a0, a1 = lo + c * jpp, min(hi, lo + (c + 1) * jpp)
    done = a0 - lo
    for grp in _merge_whole_items(job_segments(a0, a1, p), p):
      n0, n1 = grp[0][0], grp[-1][0] + 1
      p0, p1 = grp[0][1], grp[0][2]
      cnt = (n1 - n0) * (p1 - p0)
      s0 = slot_of(done)
      views = {k: b[s0:s0 + cnt] for k, b in bufs.items()}
      o = compute_fn(rgb[n0:n1], depth[n0:n1], src_pos[n0:n1], tgt_pos[n0:n1, p0:p1], depth_scale=depth_scale,
                     per_job_bin=(bin_mode == 'job'), export_bin=(bin_mode in ('call', 'shard')), out=dict(views),
                     **(dict(kwargs, compact=True) if compact else kwargs))
      if bin_mode in ('call', 'shard'):
        bins.append(o['bin'].clone())
      for k, v in views.items():  # a compute_fn that ignores `out` (the CPU stand-in) returned fresh tensors
        if o[k].data_ptr() != v.data_ptr():
          v.copy_(o[k])
      done += cnt
    if multi:  # in-place all-gather of piece c (async: it travels on NCCL's stream while the next piece renders)
      for k, b in bufs.items():
        blk = b[c * world * jpp:(c + 1) * world * jpp]
        handles.append(dist.all_gather_into_tensor(blk, blk[rank * jpp:(rank + 1) * jpp], group=group, async_op=npieces > 1))
  for hnd in handles:
    if hnd is not None:
      hnd.wait()

  # slot -> job: slot (c, r, j) holds job r * cap + c * jpp + j
  job_map = None
  if multi and npieces > 1:
    cs, rs, js = torch.meshgrid(torch.arange(npieces), torch.arange(world), torch.arange(jpp), indexing='ij')
    job_map = (rs * cap + cs * jpp + js).reshape(-1).to(torch.int32)

  # reject bin: one MAX all-reduce of (-min depth, max R, max G, max B, -own winner depth of global job 0)
  owners = []  # (slot of the owner job, 5-vector) to apply on this rank's tensors
  if bin_mode in ('call', 'shard'):
    ninf = -float('inf')
    if bins:
      b = torch.stack(bins).to(torch.float32)
      red = torch.cat([(-b[:, :1]).max(dim=0).values, b[:, 1:4].max(dim=0).values, -b[0, 4:5]])
    else:
      red = torch.tensor([ninf, 0.0, 0.0, 0.0, ninf], device=dev)
    if bin_mode == 'call':
      if not (lo == 0 and hi > 0):
        red[4] = ninf  # only the owner of global job 0 knows its pixel's own winner
      if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX, group=group)
      if jobs > 0 and (gather or (lo == 0 and hi > 0)):
        owners.append((0, torch.cat([-red[:1], red[1:4], -red[4:5]])))
    else:  # 'shard': every rank's first job owns that rank's bin
      mine = torch.cat([-red[:1], red[1:4], -red[4:5]])
      if gather and world > 1:
        allb = torch.empty((world, 5), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allb, mine.contiguous(), group=group)
        for r in range(world):
          a, b_ = shard_bounds(jobs, r, world)
          if b_ > a:
            owners.append((r * jpp if npieces > 1 else r * cap, allb[r]))
      elif hi > lo:
        owners.append((0, mine))
  for slot, vec in owners:
    tgt = {k: b[slot:] for k, b in bufs.items()}
    if raw:
      apply_bin_fn(vec, tgt, depth_scale, raw_features=True)
    else:
      apply_bin_fn(vec, tgt, depth_scale)

  # padded slots -> job order (only when the job count does not divide by the world size)
  if multi and jobs % world != 0:
    idx = torch.cat([torch.arange(r * cap, r * cap + (shard_bounds(jobs, r, world)[1] - shard_bounds(jobs, r, world)[0]))
                     for r in range(world)]).to(dev)
    bufs = {k: b.index_select(0, idx) for k, b in bufs.items()}
  result = dict(bufs)
  if compact and expand:
    result = expand_fn(result) if job_map is None else expand_fn(result, job_map=job_map.to(dev))
  result['job_range'] = (lo, hi)
  return result


class ShardedReprojection(object):
  """`reproject_sharded` with everything that does not depend on the data done once (the serving form: the same
  shapes call after call).  The gather buffers, the prepared kernel calls of every piece (one C-ABI call each), the
  slot -> job map and the bin vector are built in __init__; `run()` is: per piece one prepared call and its in-place
  all-gathers (async, they travel while the next piece renders), one 5-float all-reduce for the reject bin, the bin
  patch and -- with expand=True -- one expand kernel that writes the float32 tensors in job order.

  wire='multicast' (NVSwitch): the gather buffer is symmetric memory (torch.distributed._symmetric_memory) and the
  resolve kernel of every rank stores its compact planes through the buffer's MULTICAST mapping -- one store
  instruction lands in the same slot of every rank's buffer, the switch replicates it.  Compute and collective are
  one kernel: there is no all-gather call, no pieces, and a rank's NVLink egress is its own 7 B per pixel instead of
  (world - 1) x that.  Two symmetric-memory barriers per run() order the stores against the readers of the
  previous and of this result.  Raises RuntimeError where the platform has no multicast (callers fall back to
  wire='nccl').

  Restrictions of the prepared form: the job count divides by the world size, bin_mode 'call' or 'job', compact wire.
  The tensors passed in are kept and read in place by every run()."""

  def __init__(self, rgb, depth, src_pos, tgt_pos, *, group=None, bin_mode: str = 'call', pieces: int = 2,
               gather: bool = True, expand: bool = True, depth_scale: float = constants.DEPTH_SCALE, mask_frames: int = 0,
               wire: str = 'nccl', **conv):
    from . import guidance
    self.g = guidance
    self.group, self.bin_mode, self.gather, self.expand, self.depth_scale = group, bin_mode, gather, expand, depth_scale
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    rgb = torch.as_tensor(rgb)
    if rgb.dim() == 4:
      rgb = rgb[:, None]
    n, s, h, w, _ = rgb.shape
    tgt_pos = torch.as_tensor(tgt_pos).reshape(n, -1, 3)
    p = tgt_pos.shape[1]
    jobs = n * p
    if bin_mode not in ('call', 'job') or jobs % self.world:
      raise ValueError('the prepared form needs bin_mode call / job and a job count that divides by the world size')
    depth = torch.as_tensor(depth).reshape(n, s, h, w)
    src_pos = torch.as_tensor(src_pos).reshape(n, s, 3)
    dev = rgb.device
    world, rank = self.world, self.rank
    if wire not in ('nccl', 'multicast'):
      raise ValueError("wire must be 'nccl' or 'multicast'")
    multi = gather and world > 1
    self.mcast = multi and wire == 'multicast'
    cap = jobs // world
    lo = rank * cap
    npieces = max(1, min(cap, pieces)) if multi and not self.mcast else 1  # multicast: the transfer is inside the kernel
    while cap % npieces:
      npieces -= 1
    jpp = cap // npieces
    slots = npieces * world * jpp if multi else cap
    self.symm = None
    if self.mcast:
      import torch.distributed._symmetric_memory as symm_mem
      px = h * w
      depth_off = (slots * px * 3 + 255) // 256 * 256  # [uint8 colours | float32 depth] in one symmetric allocation
      self._symm_buf = symm_mem.empty(depth_off + slots * px * 4, dtype=torch.uint8, device=dev)
      self.symm = symm_mem.rendezvous(self._symm_buf, group if group is not None else dist.group.WORLD)
      mc = int(self.symm.multicast_ptr)
      if not mc:
        raise RuntimeError('symmetric memory on this platform has no multicast mapping (NVSwitch multicast unavailable)')
      self.rgb8 = self._symm_buf[:slots * px * 3].view(slots, h, w, 3)
      self.depth = self._symm_buf[depth_off:depth_off + slots * px * 4].view(torch.float32).view(slots, h, w, 1)
    else:
      self.rgb8 = torch.empty((slots, h, w, 3), dtype=torch.uint8, device=dev)
      self.depth = torch.empty((slots, h, w, 1), dtype=torch.float32, device=dev)
    self.multi, self.npieces, self.world_jpp, self.jpp, self.jobs, self.lo, self.hi = multi, npieces, world * jpp, jpp, jobs, lo, lo + cap
    self.calls = []   # (piece, PreparedReprojection, bin tensor or None)
    for c in range(npieces):
      a0 = lo + c * jpp
      done = 0
      for grp in _merge_whole_items(job_segments(a0, a0 + jpp, p), p):
        n0, n1 = grp[0][0], grp[-1][0] + 1
        p0, p1 = grp[0][1], grp[0][2]
        cnt = (n1 - n0) * (p1 - p0)
        s0 = ((c * world + rank) * jpp + done) if multi else c * jpp + done
        out = {'proj_rgb_u8': self.rgb8[s0:s0 + cnt], 'proj_depth': self.depth[s0:s0 + cnt]}
        binb = torch.zeros(5, device=dev) if bin_mode == 'call' else None
        plan = guidance.prepare(rgb[n0:n1].contiguous(), depth[n0:n1].contiguous(), src_pos[n0:n1].contiguous(),
                                tgt_pos[n0:n1, p0:p1].contiguous(), depth_scale=depth_scale, mask_frames=mask_frames,
                                per_job_bin=(bin_mode == 'job'), compact=True, out=out, bin_out=binb, **conv)
        if self.mcast:  # same slot, every rank's buffer
          plan.redirect_outputs(mc + s0 * px * 3, mc + depth_off + s0 * px * 4)
        self.calls.append((c, plan, binb))
        done += cnt
    self.job_map = None
    if multi and npieces > 1:
      cs, rs, js = torch.meshgrid(torch.arange(npieces), torch.arange(world), torch.arange(jpp), indexing='ij')
      self.job_map = (rs * cap + cs * jpp + js).reshape(-1).to(device=dev, dtype=torch.int32)
    self.red = torch.zeros(5, device=dev)
    self.owns_job0 = lo == 0

  def run(self) -> Dict[str, torch.Tensor]:
    handles = []
    if self.mcast:
      self.symm.barrier(channel=0)  # every rank is done reading the previous result: its slots may be overwritten
    for i, (c, plan, _) in enumerate(self.calls):
      plan.run()
      last_of_piece = i + 1 == len(self.calls) or self.calls[i + 1][0] != c
      if self.multi and last_of_piece and not self.mcast:
        for b in (self.rgb8, self.depth):
          blk = b[c * self.world_jpp:(c + 1) * self.world_jpp]
          handles.append(dist.all_gather_into_tensor(blk, blk[self.rank * self.jpp:(self.rank + 1) * self.jpp],
                                                     group=self.group, async_op=True))
    if self.bin_mode == 'call':
      bins = torch.stack([b for _, _, b in self.calls])
      red = self.red
      red[0] = (-bins[:, 0]).max()
      red[1:4] = bins[:, 1:4].max(dim=0).values
      red[4] = -bins[0, 4] if self.owns_job0 else -float('inf')
      if self.world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX, group=self.group)
    for hnd in handles:
      hnd.wait()
    if self.mcast:
      self.symm.barrier(channel=1)  # every rank's resolve kernels have finished: all slots of this buffer are complete
    out = {'proj_rgb_u8': self.rgb8, 'proj_depth': self.depth}
    if self.bin_mode == 'call' and self.jobs > 0 and (self.multi or self.owns_job0):
      vec = torch.cat([-self.red[:1], self.red[1:4], -self.red[4:5]])
      self.g.apply_bin(vec, out, self.depth_scale)
    if self.expand:
      out = self.g.expand_guidance(dict(out), job_map=self.job_map)
    elif self.job_map is not None:
      out['job_map'] = self.job_map
    out['job_range'] = (self.lo, self.hi) if not self.multi else (0, self.jobs)
    return out
