#!/usr/bin/env python
"""bench.py -- throughput of the geometric guidance path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--dist room|rand]
  python bench.py --impl reference ...     # the reference path on the host CPU cores

A "step" is one pass of the fused path over one batch of synthetic RGB-D panoramas
(config c2 = 512x1024, batch 8, one source frame, one target pose -- BASELINE.json configs[1]).
Weak scaling: every rank (one process per GPU) runs the full config on its own batch; no
data-path collective.  Prints ONE JSON line on rank 0; `extra` holds short sub-records of the other
configs / modes (64-bit key, compact outputs, c1 / c3 / c4 / c5) and, on several GPUs, the strong-scaling
sharded form of c4 with and without the NCCL all-gather of the guidance.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    'c1': dict(n=1, s=1, p=1, h=256, sweep=False),   # lowres, batch 1 (the reference's CPU-runnable case)
    'c2': dict(n=8, s=1, p=1, h=512, sweep=False),   # highres, batch 8  <- the metric's config
    'c3': dict(n=32, s=4, p=1, h=512, sweep=False),  # trajectory accumulation
    'c4': dict(n=1, s=1, p=64, h=512, sweep=True),   # VLN perturbation sweep
    'c5': dict(n=64, s=8, p=1, h=2048, sweep=False), # stress
}
L2_BYTES = 126 << 20


def alg_bytes(c):
  """7 B per unique source point + 20 B per target pixel."""
  hw = c['h'] * 2 * c['h']
  return c['n'] * c['s'] * hw * 7, c['n'] * c['p'] * hw * 20


def measured_peak_gbs():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      return float(json.load(f)['hbm_gbs']), 'measured'
  except Exception:  # pylint: disable=broad-except
    return 6650.0, 'fallback'


def committed_traffic():
  """Per-launch DRAM bytes of the dominant kernel from the committed ncu summary, if any."""
  try:
    with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
      return json.load(f)
  except Exception:  # pylint: disable=broad-except
    return None


class ClockSampler(threading.Thread):
  """Samples SM clock and throttle reasons through NVML while the timed region runs."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
    self._stop_evt = threading.Event()
    self._armed = threading.Event()  # set right before the timed region: the NVML set-up and the thread start happen earlier
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
      pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)  # warm the NVML path up before the timed region
    except Exception:  # pylint: disable=broad-except
      self.nv = None

  def run(self):
    if self.nv is None:
      return
    nv = self.nv
    names = {
        'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8)),
        'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40)),
        'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20)),
        'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)),
    }
    get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
    while not self._stop_evt.is_set():
      if not self._armed.is_set():
        time.sleep(0.0002)
        continue
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = get_reasons(self.h)
        for k, bit in names.items():
          if r & bit:
            self.reasons.add(k)
      except Exception:  # pylint: disable=broad-except
        pass
      time.sleep(0.001)

  def arm(self):
    self._armed.set()

  def stop(self):
    self._stop_evt.set()
    self.join(timeout=2)
    return {'sm_mhz': statistics.median(self.samples) if self.samples else None, 'sm_max_mhz': self.max_mhz,
            'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/): the restated reference path on host cores
# ------------------------------------------------------------------------------------------
def _cpu_one_batch(args):
  """One reference-style call on `n` panoramas: literal numpy restatement (TF absent)."""
  from oracle import ref_numpy as R
  inp, which = args
  t0 = time.perf_counter()
  if which == 'numpy':
    R.reproject_trajectory(inp['rgb'], inp['depth'], inp['src_pos'], inp['tgt_pos'][:, 0], mask_first_frame=True)
  else:
    from oracle import ref_exact as X
    X.reproject(inp['rgb'], inp['depth'], inp['src_pos'], inp['tgt_pos'], mask_first_frame=True)
  return time.perf_counter() - t0


def cpu_baseline_single(cfg, dist, budget_s=12.0):
  """cpu_baseline of our arm: 1 host core, a bounded sample of the same workload."""
  from se3ds_b200 import synth
  c = dict(cfg)
  inp = synth.make_inputs(1, c['s'], c['p'], c['h'], seed=100, dist=dist, sweep=c['sweep'])
  times = {}
  for which in ('numpy', 'c'):
    _cpu_one_batch((inp, which))  # warm-up (also builds / loads the C oracle)
    ts, t_start = [], time.perf_counter()
    while len(ts) < 3 or (time.perf_counter() - t_start < budget_s / 2 and len(ts) < 40):
      ts.append(_cpu_one_batch((inp, which)))
    times[which] = statistics.median(ts)
  best = min(times, key=times.get)
  panos = c['p']  # one item, P target panoramas
  return {'value': panos / times[best], 'unit': 'panos/s', 'cores': 1, 'kind': 'port',
          'sample': (f"1 item of the workload ({c['s']} frame(s), {c['p']} pose(s), {c['h']}x{2 * c['h']}) per call, "
                     f"median of repeated calls; numpy restatement {times['numpy'] * 1e3:.1f} ms, "
                     f"scalar C restatement {times['c'] * 1e3:.1f} ms per call; fastest ({best}) reported; "
                     'TensorFlow is not installed, so this is the restated reference path'),
          'mpoints_per_s': c['s'] * c['p'] * c['h'] * 2 * c['h'] / times[best] / 1e6}


def run_reference_arm(args, cfg):
  """--impl reference: the restated reference path on all host cores (process pool, one item per task)."""
  import multiprocessing as mp
  from se3ds_b200 import synth
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count()
  workers = max(1, min(cores, 64))
  c = dict(cfg)
  inp = synth.make_inputs(1, c['s'], c['p'], c['h'], seed=100, dist=args.dist, sweep=c['sweep'])
  # one step = `workers` items (a bounded sample of the config's batch), one per worker
  steps, warmup = max(1, args.steps), max(1, args.warmup)
  budget_s = 150.0  # the whole arm has to end within a few minutes: fewer steps than asked only beyond this
  which = 'c'
  capped = None
  with mp.get_context('fork').Pool(workers) as pool:
    t0 = time.perf_counter()
    for _ in range(warmup):
      pool.map(_cpu_one_batch, [(inp, which)] * workers)
    per_step = (time.perf_counter() - t0) / warmup
    if per_step * steps > budget_s:
      capped = f'{steps} steps asked; {per_step:.2f} s per step would exceed the {budget_s:.0f} s budget of the CPU arm'
      steps = max(1, int(budget_s / per_step))
    t0 = time.perf_counter()
    for _ in range(steps):
      pool.map(_cpu_one_batch, [(inp, which)] * workers)
    dt = time.perf_counter() - t0
  panos = workers * c['p'] * steps
  value = panos / dt
  line = {
      'impl': 'reference', 'metric': 'reprojected panoramas/s', 'value': value, 'unit': 'panos/s', 'n_gpus': args.gpus,
      'steps': steps, 'warmup': warmup, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(args.config, c, args.dist)},
      'sample_items_per_step': workers, 'steps_capped': capped,
      'cpu_baseline': {'value': value, 'unit': 'panos/s', 'cores': workers, 'kind': 'port',
                       'sample': f'{workers} items per step (one per worker process), scalar C restatement of the '
                                 'reference path (TensorFlow not installed)'},
      'e2e': {'value': value, 'unit': 'panos/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'mpoints_per_s': value * c['s'] * c['h'] * 2 * c['h'] / 1e6,
  }
  print(json.dumps(line), flush=True)


def workload_name(name, c, dist):
  return (f"{name}: N{c['n']} S{c['s']} P{c['p']} {c['h']}x{2 * c['h']} equirect RGB-D (u8 RGB + f32 depth), "
          f"depth={dist}, mask frame 0, void -1/-1")


def sharded_record(args, world, rank, dev, steps=60, warmup=5):
  """Strong scaling of ONE call (config c4: 64 poses of one pano): every rank renders its block of the job list;
  `gather`: the finished guidance is all-gathered to every rank inside the step (compact wire format, in place,
  pipelined, expanded to float32 on arrival -- se3ds_b200/parallel.py); `local`: the shards stay where they
  were rendered (no data-path collective).  Returns the sub-record (max over ranks)."""
  import torch
  import torch.distributed as dist
  from se3ds_b200 import parallel, synth
  cfg = CONFIGS['c4']
  n, s, p, h = cfg['n'], cfg['s'], cfg['p'], cfg['h']
  inp = synth.make_inputs(n, s, p, h, seed=7, dist=args.dist, sweep=True)
  t = {k: torch.as_tensor(v).to(dev) for k, v in inp.items()}
  rec = {'workload': workload_name('c4', cfg, args.dist), 'scaling': 'strong', 'steps': steps}
  for label, kw in (('gather_compact', dict(gather=True, wire='compact')), ('gather_f32', dict(gather=True, wire='f32')),
                    ('local', dict(gather=False))):
    def step():
      return parallel.reproject_sharded(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, **kw)
    for _ in range(warmup):
      out = step()
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec[label] = {'ms_per_step': ms.item(), 'panos_per_s': n * p / (ms.item() * 1e-3),
                  'bytes_held_per_rank': sum(v.numel() * v.element_size() for v in out.values() if torch.is_tensor(v))}
  # the prepared (serving) form: buffers, kernel calls and maps built once, run() repeated
  for label, kw in (('prepared_gather_expand', dict(gather=True, expand=True)), ('prepared_gather_compact', dict(gather=True, expand=False)),
                    ('prepared_multicast_expand', dict(gather=True, expand=True, wire='multicast')),
                    ('prepared_multicast_compact', dict(gather=True, expand=False, wire='multicast')),
                    ('prepared_local', dict(gather=False, expand=False))):
    plan, why = None, ''
    try:
      plan = parallel.ShardedReprojection(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, pieces=args.pieces, **kw)
    except Exception as e:  # pylint: disable=broad-except
      if kw.get('wire') != 'multicast':
        raise
      why = f'{type(e).__name__}: {e}'[:200]
    if kw.get('wire') == 'multicast':  # all ranks take the multicast path or none does
      okf = torch.tensor([0 if plan is None else 1], device=dev)
      if world > 1:
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
      if okf.item() == 0:
        rec[label] = {'unavailable': why or 'another rank could not set up the multicast mapping'}
        del plan
        continue
    for _ in range(warmup):
      out = plan.run()
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      out = plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec[label] = {'ms_per_step': ms.item(), 'panos_per_s': n * p / (ms.item() * 1e-3), 'pieces': plan.npieces,
                  'bytes_held_per_rank': sum(v.numel() * v.element_size() for v in out.values() if torch.is_tensor(v))}
    del plan
  return rec


def compat_record(torch, args, dev, steps=100):
  """The un-fused drop-in calls of the reference signatures on a c2-sized MATERIALISED cloud (what models/models.py:
  211-226,276-281 executes without code changes): pano_utils.equirectangular_to_pointcloud, then
  pano_utils.project_feats_to_equirectangular on (xyz1 + src) - tgt.  Each call is timed by itself with CUDA events;
  the two translations in between are the caller's elementwise glue and are not timed."""
  from se3ds_b200 import synth
  from se3ds_b200.utils import pano_utils
  cfg = CONFIGS['c2']
  n, h = cfg['n'], cfg['h']
  w, hw = 2 * h, 2 * h * h
  peak = measured_peak_gbs()[0]
  sets = []
  for r in range(3):  # 3 x (16 + 28 + 16) B per point and pixel > 2x L2
    inp = synth.make_inputs(n, 1, 1, h, seed=500 + r, dist=args.dist)
    rgb = torch.as_tensor(inp['rgb'][:, 0].astype(np.int32)).to(dev)
    depth = torch.as_tensor(inp['depth'][:, 0]).to(dev)
    off = torch.as_tensor(np.concatenate([inp['src_pos'][:, 0] - inp['tgt_pos'][:, 0], np.zeros((n, 1), np.float32)], 1)).to(dev)
    xyz1, feats = pano_utils.equirectangular_to_pointcloud(rgb, depth, -1, 20.0)
    sets.append((rgb, depth, (xyz1 + off[:, :, None]).contiguous(), feats))
  def timed(fn):
    for i in range(6):
      fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
      fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
  t_un = timed(lambda i: pano_utils.equirectangular_to_pointcloud(sets[i % 3][0], sets[i % 3][1], -1, 20.0))
  t_pr = timed(lambda i: pano_utils.project_feats_to_equirectangular(sets[i % 3][3], sets[i % 3][2], h, w, -1, 20.0))
  un_bytes = n * hw * (4 + 12 + 16 + 12)  # f32 depth + int32 colours in, xyz1 + int32 features out
  pr_bytes = n * hw * (16 + 12) + n * hw * (4 + 12)  # xyz1 + int32 features in, f32 depth + f32 features out
  return {'workload': 'c2-sized materialised cloud: N8 512x1024, int32 colours (void -1), reference signatures, device tensors',
          'unproject': {'ms': t_un, 'alg_bytes': un_bytes, 'frac_of_hbm_peak': un_bytes / (t_un * 1e-3) / 1e9 / peak},
          'project': {'ms': t_pr, 'alg_bytes': pr_bytes, 'frac_of_hbm_peak': pr_bytes / (t_pr * 1e-3) / 1e9 / peak,
                      'panos_per_s': n / (t_pr * 1e-3)},
          'note': 'what the reference callers get without code changes; the fused path above replaces both calls and the glue'}


def gpu_measure(guidance, synth, lib, torch, dist, cfg, args, dev, local_rank, world, rank, steps, warmup,
                key64=False, compact=False, clocks=False):
  """Device-resident throughput of one config: K timed steps of the fused path over a ring of input / output sets
  larger than 2x L2, stream launches with programmatic dependent launch, CUDA events, max over ranks."""
  n, s, p, h = cfg['n'], cfg['s'], cfg['p'], cfg['h']
  src_bytes, out_bytes = alg_bytes(cfg)
  set_bytes = src_bytes + out_bytes
  ring = max(2, min(16, -(-2 * L2_BYTES // set_bytes) + 1))
  ws = lib.Workspace(local_rank, 0, args.chunk_mb << 20)
  ws.projection_mode(args.proj_mode)
  ws.pdl(not args.no_pdl)
  if args.lanes:
    ws.lanes(args.lanes, 0, args.lane_chunks)
  plans = []
  for r in range(ring):
    gen_n = min(n, 8)  # large configs reuse one generated block of items per set
    inp = synth.make_inputs(gen_n, s, p, h, seed=1000 * rank + r, dist=args.dist, sweep=cfg['sweep'])
    if gen_n < n:
      inp = {k: np.concatenate([v] * (n // gen_n), axis=0) for k, v in inp.items()}
    t = {k: torch.as_tensor(v).to(dev) for k, v in inp.items()}
    plans.append(guidance.prepare(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, workspace=ws,
                                  key64=key64, inputs_ready=True, compact=compact))
  stream = torch.cuda.Stream(dev)

  def barrier():
    stream.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  with torch.cuda.stream(stream):
    for pl in plans:  # grows the workspace, uploads the tables
      pl.run()
    stream.synchronize()
    graphs = None
    if args.graph:
      graphs = []
      for pl in plans:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
          pl.run()
        graphs.append(g)

    def step(i):
      if graphs is not None:
        graphs[i % ring].replay()
      else:
        plans[i % ring].run()

    l0 = ws.profile_read()[1]
    plans[0].run()
    stream.synchronize()
    launches_per_step = ws.profile_read()[1] - l0
    # the sampler's NVML set-up (milliseconds) and its thread start come BEFORE the warm-up, so that nothing but the
    # contract's barrier + synchronize sits between the last warm-up step and the first timed one (a GPU left idle for
    # milliseconds starts the timed burst from a lower power state: +5 % on a 20-step run)
    late = bool(os.environ.get('SE3DS_BENCH_LATE_SAMPLER'))
    sampler = ClockSampler(local_rank) if clocks and not late else None
    if sampler:
      sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warmup):
      step(i)
    barrier()
    if clocks and late:
      sampler = ClockSampler(local_rank)
      sampler.start()
    if sampler:
      sampler.arm()
    e0.record(stream)
    for i in range(steps):
      step(i)
    e1.record(stream)
    barrier()
    clk = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)

    # the kernels' shares of the pipelined step: end-of-kernel stamps, programmatic dependent launch on
    prof_steps = min(steps, 300)
    ws.profile(2)
    for i in range(prof_steps + 1):
      plans[i % ring].run()
    shares, chunks = ws.profile_read_stamps()
    ws.profile(0)
    steps_counted = max(1, chunks) / max(1, launches_per_step // 3)
    shares = [x / steps_counted for x in shares]
    # ... and their durations run one at a time (cudaEvents between the launches, no overlap)
    ws.profile(1)
    for i in range(prof_steps):
      plans[i % ring].run()
    serial, _ = ws.profile_read()
    ws.profile(0)
    serial = [x / prof_steps for x in serial]
  t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_step = float(t.item()) / steps
  del plans
  ws.close()
  return dict(ms_step=ms_step, launches_per_step=launches_per_step, clocks=clk, shares=shares, serial=serial, ring=ring,
              set_bytes=set_bytes, src_bytes=src_bytes, out_bytes=out_bytes, graph=graphs is not None)


def e2e_measure(guidance, torch, dist, host_inp, args, local_rank, world, compact, ws, pipelined=False):
  """The same metric through the reference-facing call with HOST buffers: se3ds_reproject_host copies the inputs
  from pinned memory, runs the kernels and copies the guidance back, all inside the timed region.  pipelined:
  guidance.HostReprojector (two workspaces, SE3DS_FLAG_HOST_ASYNC) -- step i+1 is submitted before step i is
  waited for, so its upload overlaps the download of step i; every step's result is complete in host memory
  before the clock stops."""
  out = {}
  kw = dict(mask_frames=1, device=local_rank, compact=compact)
  batch = (host_inp['rgb'], host_inp['depth'], host_inp['src_pos'], host_inp['tgt_pos'])
  pipe = guidance.HostReprojector(depth=args.e2e_depth, **kw) if pipelined else None
  def steps(k):
    nonlocal out
    if pipelined:
      done = 0
      for _ in range(k):
        r = pipe.submit(*batch)
        if r is not None:
          out, done = r, done + 1
      for r in pipe.flush():
        out, done = r, done + 1
      assert done == k
    else:
      for _ in range(k):
        guidance.reproject_host(*batch, out=out, workspace=ws, **kw)
  steps(args.e2e_depth + 2 if pipelined else 3)  # every rotating output set is allocated (pinned) before the clock starts
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  steps(args.e2e_steps)
  torch.cuda.synchronize()
  ms = (time.perf_counter() - t0) / args.e2e_steps * 1e3
  t = torch.tensor([ms], device=torch.device('cuda', local_rank), dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  h2d = sum(host_inp[k].numel() * host_inp[k].element_size() for k in host_inp)
  d2h = sum(v.numel() * v.element_size() for v in out.values())
  if pipe is not None:
    pipe.close()
  return float(t.item()), h2d, d2h, out


def copy_ceiling(torch, dev, h2d_bytes, d2h_bytes, reps=10):
  """What the host link gives this rank for the same byte counts: pinned H2D and D2H copies alone, each timed by itself."""
  hb = torch.empty(max(h2d_bytes, d2h_bytes), dtype=torch.uint8, pin_memory=True)
  db = torch.empty(max(h2d_bytes, d2h_bytes), dtype=torch.uint8, device=dev)
  res = {}
  for name, nbytes, fn in (('h2d_gbs', h2d_bytes, lambda: db[:h2d_bytes].copy_(hb[:h2d_bytes], non_blocking=True)),
                           ('d2h_gbs', d2h_bytes, lambda: hb[:d2h_bytes].copy_(db[:d2h_bytes], non_blocking=True))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
      fn()
    torch.cuda.synchronize()
    res[name] = nbytes * reps / (time.perf_counter() - t0) / 1e9
  return res


def parity_vs_libm(guidance, torch, inp, dev):
  """GPU result of one item of the workload against the literal numpy (libm) restatement: the two float32
  pipelines differ only where an ulp of a transcendental moves a point across a pixel border
  (tests/test_oracle_disagreement.py).  Part of the CPU-baseline leg (the only place bench.py runs oracle/)."""
  from oracle import ref_numpy as R
  one = {k: v[:1] for k, v in inp.items()}
  t = {k: torch.as_tensor(v).to(dev) for k, v in one.items()}
  out = guidance.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1)
  torch.cuda.synchronize()
  image, depth, mask, _ = R.reproject_trajectory(one['rgb'], one['depth'], one['src_pos'], one['tgt_pos'][:, 0], mask_first_frame=True)
  gd, gi, gm = (out[k].cpu().numpy() for k in ('proj_depth', 'proj_image', 'proj_mask'))
  bad_d = np.abs(gd - depth) > 1e-5 * np.abs(depth) + 1e-7
  return {'pixels': int(depth.size), 'depth_differs': float(bad_d.mean()), 'rgb_differs': float(np.any(gi != image, axis=-1).mean()),
          'mask_differs': float((gm != mask).mean()),
          'note': 'GPU (canonical float32 arithmetic) vs numpy restatement (libm); differences sit on pixel borders'}


# ------------------------------------------------------------------------------------------
def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=2000)
  ap.add_argument('--warmup', type=int, default=20)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
  ap.add_argument('--dist', default='room', choices=['room', 'rand'])
  ap.add_argument('--graph', action='store_true', help='CUDA-graph replay instead of stream launches (measured slower: '
                  'stream launches keep the programmatic dependent launch overlap)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-extras', action='store_true', help='skip the sub-records of the other configs / modes')
  ap.add_argument('--e2e-steps', type=int, default=20, help='0 skips the host-buffer leg (very large configs: it pins the whole input set)')
  ap.add_argument('--e2e-depth', type=int, default=2, help='workspaces in flight in the pipelined host leg (guidance.HostReprojector)')
  ap.add_argument('--chunk-mb', type=int, default=0, help='workspace L2 chunk size (0 = library default)')
  ap.add_argument('--n-override', type=int, default=0, help='override the batch size of the config (memory-bounded runs)')
  ap.add_argument('--key64', action='store_true', help='force the 64-bit packed depth|index z-buffer key')
  ap.add_argument('--lanes', type=int, default=0, help='concurrent chunk lanes of one call (0 = library default, 2)')
  ap.add_argument('--lane-chunks', type=int, default=0, help='minimum chunks per lane (0 = library default, 2)')
  ap.add_argument('--no-pdl', action='store_true', help='disable programmatic dependent launch')
  ap.add_argument('--pieces', type=int, default=2, help='pieces of the pipelined gather in the sharded c4 sub-record')
  ap.add_argument('--proj-mode', type=int, default=1, help='0 canonical projection only, 1 certified fast path (default)')
  args = ap.parse_args()
  cfg = dict(CONFIGS[args.config])
  if args.n_override:
    cfg['n'] = args.n_override
  if args.impl == 'reference':
    run_reference_arm(args, cfg)
    return
  args.warmup = max(args.warmup, 3)

  import torch
  import torch.distributed as dist
  from se3ds_b200 import _lib, guidance, hostmem, synth

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  numa = hostmem.bind_to_gpu_numa_node(local_rank)  # pinned buffers are placed by first touch: allocate them next to the GPU
  dev = torch.device('cuda', local_rank)
  if numa['cpus'] is None:  # sysfs did not say: let NVML pick the CPUs next to this rank's GPU
    try:
      import pynvml
      pynvml.nvmlInit()
      pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
      numa['nvml_affinity'] = len(os.sched_getaffinity(0))
    except Exception:  # pylint: disable=broad-except
      pass
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)

  n, s, p, h = cfg['n'], cfg['s'], cfg['p'], cfg['h']
  w = 2 * h
  mods = (guidance, synth, _lib, torch, dist)
  m = gpu_measure(*mods, cfg, args, dev, local_rank, world, rank, args.steps, args.warmup, key64=args.key64, clocks=True)
  ms_step = m['ms_step']
  src_bytes, out_bytes = m['src_bytes'], m['out_bytes']
  panos_per_s = world * n * p / (ms_step * 1e-3)
  mpoints = world * n * s * p * h * w / (ms_step * 1e-3) / 1e6

  # e2e: host buffers, H2D + kernels + D2H inside the C-ABI call (se3ds_reproject_host)
  gen_n = min(n, 8)
  inp = synth.make_inputs(gen_n, s, p, h, seed=1000 * rank + 99, dist=args.dist, sweep=cfg['sweep'])
  if gen_n < n:
    inp = {k: np.concatenate([v] * (n // gen_n), axis=0) for k, v in inp.items()}
  if args.e2e_steps > 0:
    host_inp = {k: torch.as_tensor(v).pin_memory() for k, v in inp.items()}
    ews = _lib.Workspace(local_rank, 0, args.chunk_mb << 20)
    e2e_ms, h2d, d2h, _ = e2e_measure(guidance, torch, dist, host_inp, args, local_rank, world, False, ews)
    e2e_c_ms, h2d_c, d2h_c, _ = e2e_measure(guidance, torch, dist, host_inp, args, local_rank, world, True, ews)
    e2e_p_ms = e2e_measure(guidance, torch, dist, host_inp, args, local_rank, world, False, None, pipelined=True)[0]
    e2e_pc_ms = e2e_measure(guidance, torch, dist, host_inp, args, local_rank, world, True, None, pipelined=True)[0]
    ews.close()
    ceiling = copy_ceiling(torch, dev, h2d, d2h)
  else:
    e2e_ms = e2e_c_ms = e2e_p_ms = e2e_pc_ms = float('inf')
    h2d = d2h = h2d_c = d2h_c = 0
    ceiling = None

  extras = {}
  if not args.no_extras:
    # other configs / modes, short runs (sub-records; the headline stays the config above)
    def sub(name, c, **kw):
      nsteps = kw.pop('steps')
      r = gpu_measure(*mods, c, args, dev, local_rank, world, rank, nsteps, 5, **kw)
      sb, ob = alg_bytes(c)
      hw = c['h'] * 2 * c['h']
      if min(r['shares']) < 0:  # several chunks on two lanes: a chunk's first kernel ends before the other lane's last one
        r['shares'] = None
      return {'workload': workload_name(name, c, args.dist), 'ms_per_step': r['ms_step'], 'steps': nsteps,
              'panos_per_s': world * c['n'] * c['p'] / (r['ms_step'] * 1e-3),
              'mpoints_per_s': world * c['n'] * c['s'] * c['p'] * hw / (r['ms_step'] * 1e-3) / 1e6,
              'roofline_step_frac': (sb + ob) / (r['ms_step'] * 1e-3) / 1e9 / measured_peak_gbs()[0],
              'kernel_shares_ms': r['shares']}
    if args.config == 'c2' and not args.key64:
      extras['c2_key64'] = sub('c2', CONFIGS['c2'], steps=300, key64=True)
      extras['c2_key64']['note'] = '64-bit packed depth|point-index z-buffer key (what return_winner uses)'
      extras['c2_compact'] = sub('c2', CONFIGS['c2'], steps=300, compact=True)
      extras['c2_compact']['note'] = 'SE3DS_FLAG_COMPACT_OUT: uint8 colours + float32 depth leave the resolve kernel (7 B instead of 20 B per pixel)'
    if args.config == 'c2':
      extras['c1'] = sub('c1', CONFIGS['c1'], steps=300)
      extras['c3'] = sub('c3', CONFIGS['c3'], steps=30)
      extras['c4'] = sub('c4', CONFIGS['c4'], steps=60)
      c5 = dict(CONFIGS['c5'], n=4)
      extras['c5_n4'] = sub('c5 (batch reduced from 64 to 4 per GPU)', c5, steps=6)
    if args.config == 'c2' and world == 1:
      extras['compat_c2'] = compat_record(torch, args, dev)
    if world > 1:
      extras['sharded_c4'] = sharded_record(args, world, rank, dev)

  if rank == 0:
    peak, peak_kind = measured_peak_gbs()
    names = ['splat_depth_kernel', 'splat_feat_kernel', 'resolve_kernel']
    # the algorithmic bytes of SURVEY 8(d), credited to the kernel that moves them: K2 reads the depth plane
    # (4 B per source point), K3 the colours (3 B per source point), K4 writes the guidance (20 B per pixel)
    kalg = [src_bytes * 4 // 7, src_bytes * 3 // 7, out_bytes]
    kms = m['shares'] if sum(m['shares']) > 0 else m['serial']
    dom = max(range(3), key=lambda i: kms[i])
    traffic = committed_traffic()
    line = {
        'metric': 'reprojected panoramas/s', 'value': panos_per_s, 'unit': 'panos/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.config, cfg, args.dist), 'per_gpu_batch': n,
                   'cache': f"inputs+outputs rotate over a ring of {m['ring']} sets ({m['ring'] * m['set_bytes'] >> 20} MiB > 2x L2)",
                   'launch': 'cuda_graph_replay' if m['graph'] else 'stream launches (programmatic dependent launch)',
                   'parallelism': f'dp{world}', 'chunk_mb': args.chunk_mb or 'default', 'lanes': args.lanes or 'default',
                   'inputs_ready_flag': True, 'pdl': not args.no_pdl,
                   'zbuffer_key': 'u64 depth|index' if args.key64 else 'u32 depth (no winner index requested)',
                   'projection': 'certified_fast+canonical_fallback' if args.proj_mode == 1 else 'canonical'},
        'mpoints_per_s': mpoints,
        'e2e': {'value': world * n * p / (e2e_p_ms * 1e-3), 'unit': 'panos/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_p_ms,
                'how': f'guidance.HostReprojector: se3ds_reproject_host with SE3DS_FLAG_HOST_ASYNC on {args.e2e_depth} workspaces used in turn '
                       '(step i+1 uploads while step i downloads); every step copies its inputs from pinned host memory and its '
                       'float32 guidance tensors back, all results complete before the clock stops',
                'per_rank_gbs': {'h2d': h2d / (e2e_p_ms * 1e-3) / 1e9, 'd2h': d2h / (e2e_p_ms * 1e-3) / 1e9,
                                 'note': 'bytes of the step / time of the step (uploads, kernels and downloads overlap)'},
                'host_link_ceiling_gbs': ceiling, 'numa': numa,
                'blocking_call': {'value': world * n * p / (e2e_ms * 1e-3), 'unit': 'panos/s', 'ms_per_step': e2e_ms,
                                  'note': 'one se3ds_reproject_host call at a time (pipelined over the batch items inside the call only)'},
                'compact_out': {'value': world * n * p / (e2e_pc_ms * 1e-3), 'unit': 'panos/s', 'ms_per_step': e2e_pc_ms,
                                'h2d_bytes_per_step': h2d_c, 'd2h_bytes_per_step': d2h_c,
                                'blocking_call': {'value': world * n * p / (e2e_c_ms * 1e-3), 'unit': 'panos/s', 'ms_per_step': e2e_c_ms},
                                'note': 'opt-in SE3DS_FLAG_COMPACT_OUT: uint8 colours + float32 depth come back (mask = 0 < depth < 1); '
                                        'se3ds_expand_guidance restores the float32 tensors bit for bit'}},
        'gpu_launches': m['launches_per_step'] * args.steps,
        'clocks': m['clocks'],
        # achieved = B_alg / t with B_alg = 7 B per source point + 20 B per target pixel of the pass and t the
        # time of the pass.  The pass is three kernels of nearly equal length chained by programmatic dependent launch,
        # so the object describes the pass; the kernel that takes the largest share is in dominant_kernel.
        'roofline': {'bound': 'hbm', 'kernel': 'fused pass: splat_depth_kernel -> splat_feat_kernel -> resolve_kernel (programmatic dependent launch)',
                     'achieved': (src_bytes + out_bytes) / (ms_step * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                     'frac': (src_bytes + out_bytes) / (ms_step * 1e-3) / 1e9 / peak, 'peak_kind': peak_kind,
                     'traffic': (sum(traffic[k] for k in names) if traffic and all(k in traffic for k in names) else None),
                     'ms': ms_step, 'alg_bytes': src_bytes + out_bytes,
                     'dominant_kernel': {'name': names[dom], 'ms': kms[dom], 'alg_bytes': kalg[dom],
                                         'achieved': kalg[dom] / (kms[dom] * 1e-3) / 1e9, 'frac': kalg[dom] / (kms[dom] * 1e-3) / 1e9 / peak,
                                         'traffic': (traffic or {}).get(names[dom])},
                     'per_kernel_frac': {names[i]: kalg[i] / (kms[i] * 1e-3) / 1e9 / peak for i in range(3)},
                     'how': 'pass: algorithmic bytes of the step / CUDA-event time of the step (the timed region).  Kernels: share of the '
                            'pipelined step = last end-of-kernel %globaltimer stamp minus that of the kernel before it (se3ds_ws_profile '
                            'mode 2; the shares add up to the step), algorithmic bytes credited to the kernel that moves them: depth (4 B/pt) '
                            'to splat_depth, colours (3 B/pt) to splat_feat, guidance (20 B/px) to resolve.  traffic: cold-cache ncu DRAM bytes '
                            'of the three launches (profiles/traffic.json); warm: profiles/r02_ncu_warm_summary.txt',
                     'note': 'splat_depth (instruction issue + latency) and splat_feat (latency of a dependent gather + reduction per point) '
                             'are not HBM-bound (ncu: profiles/r02_*); resolve is the HBM-bound kernel, see roofline_hbm_kernel'},
        'roofline_hbm_kernel': {'bound': 'hbm', 'kernel': names[2], 'achieved': kalg[2] / (kms[2] * 1e-3) / 1e9,
                                'peak': peak, 'unit': 'GB/s', 'frac': kalg[2] / (kms[2] * 1e-3) / 1e9 / peak,
                                'traffic': (traffic or {}).get(names[2]), 'ms': kms[2]},
        'roofline_step': {'bound': 'hbm', 'achieved': (src_bytes + out_bytes) / (ms_step * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                          'frac': (src_bytes + out_bytes) / (ms_step * 1e-3) / 1e9 / peak, 'alg_bytes': src_bytes + out_bytes},
        'kernels': [{'name': names[i], 'ms': kms[i], 'ms_alone': m['serial'][i], 'alg_bytes': kalg[i]} for i in range(3)],
        'extra': extras,
    }
    if world == 1 and not args.no_cpu_baseline:
      line['cpu_baseline'] = cpu_baseline_single(cfg, args.dist)
      line['parity_vs_libm'] = parity_vs_libm(guidance, torch, inp, dev)
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
