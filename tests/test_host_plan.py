"""Host logic of the fused path that needs no GPU: the chunk / lane plan of a se3ds_reproject call
(se3ds_plan_chunks, the arithmetic reproject_core uses)."""
import math

import numpy as np
import pytest

from se3ds_b200 import _lib

MIB = 1 << 20


def job_bytes(s, h):
  return h * 2 * h * (16 + 8 * s)  # z-buffer (<= 8) + feature buffer (8) + scratch (8 per frame), per target pixel


def check_plan(n, s, p, h, budget, lanes, min_pts, min_chunks):
  pl = _lib.plan_chunks(n, s, p, h, budget, lanes, min_pts, min_chunks)
  ipc, pc = pl['items_per_chunk'], pl['poses_per_chunk']
  assert 1 <= pl['lanes'] <= lanes
  assert 1 <= ipc <= n and 1 <= pc <= p
  assert pc == p or ipc == 1                       # whole items with all their poses, or a block of poses of one item
  assert pl['chunk_jobs'] == ipc * pc
  assert pl['nchunks'] == math.ceil(n / ipc) * math.ceil(p / pc)     # the chunks tile the (item, pose) grid
  assert pl['chunk_jobs'] * s <= 65535 or pl['chunk_jobs'] == 1      # grid.z limit of the splat kernels
  eff_budget = (budget or 112 * MIB) // pl['lanes']
  assert pl['chunk_jobs'] * job_bytes(s, h) <= eff_budget or pl['chunk_jobs'] == 1   # L2 budget, shared by the lanes
  if pl['lanes'] > 1:
    assert pl['nchunks'] >= pl['lanes'] * (min_chunks or 2)
    assert pl['chunk_jobs'] * s * h * 2 * h >= (min_pts or MIB)
  return pl


def test_plan_of_the_baseline_configs():
  c1 = check_plan(1, 1, 1, 256, 0, 2, 0, 0)
  assert (c1['lanes'], c1['nchunks']) == (1, 1)
  c2 = check_plan(8, 1, 1, 512, 0, 2, 0, 0)          # configs[1]: one chunk, no fork / join
  assert (c2['lanes'], c2['nchunks'], c2['chunk_jobs']) == (1, 1, 8)
  c3 = check_plan(32, 4, 1, 512, 0, 2, 0, 0)
  assert c3['lanes'] == 2 and c3['nchunks'] == 16
  c4 = check_plan(1, 1, 64, 512, 0, 2, 0, 0)         # 64 poses of one pano: blocks of poses
  assert c4['lanes'] == 2 and c4['items_per_chunk'] == 1 and c4['poses_per_chunk'] * c4['nchunks'] >= 64
  c5 = check_plan(64, 8, 1, 2048, 0, 2, 0, 0)        # one 8-frame 2048x4096 job exceeds the budget: a chunk per job
  assert c5['lanes'] == 2 and c5['chunk_jobs'] == 1 and c5['nchunks'] == 64


def test_plan_one_lane_uses_the_whole_budget():
  one = check_plan(32, 4, 1, 512, 0, 1, 0, 0)
  two = check_plan(32, 4, 1, 512, 0, 2, 0, 0)
  assert one['lanes'] == 1 and one['chunk_jobs'] >= 2 * two['chunk_jobs']


def test_plan_falls_back_when_lanes_cannot_be_fed():
  # two jobs cannot give two lanes two chunks each
  assert check_plan(2, 1, 1, 512, 0, 2, 0, 0)['lanes'] == 1
  # ... unless one chunk per lane is allowed (and half a megapoint is accepted as a lane's share)
  assert check_plan(2, 1, 1, 512, 0, 2, 0, 1)['lanes'] == 1
  assert check_plan(2, 1, 1, 512, 0, 2, 1 << 19, 1)['lanes'] == 2
  # tiny panos: below the minimum lane size
  assert check_plan(64, 1, 1, 16, 1, 4, 0, 1)['lanes'] == 1
  assert check_plan(64, 1, 1, 16, 1, 4, 1, 1)['lanes'] == 4


def test_plan_properties_random():
  rng = np.random.default_rng(0)
  for _ in range(400):
    n, s, p = int(rng.integers(1, 40)), int(rng.integers(1, 9)), int(rng.integers(1, 70))
    h = int(rng.choice([3, 8, 16, 64, 256, 512, 1024]))
    budget = int(rng.choice([0, 1, 4 * MIB, 64 * MIB, 512 * MIB]))
    lanes = int(rng.integers(1, 5))
    min_pts = int(rng.choice([0, 1, 1 << 16]))
    min_chunks = int(rng.integers(0, 4))
    check_plan(n, s, p, h, budget, lanes, min_pts, min_chunks)


def test_plan_argument_errors():
  with pytest.raises(ValueError):
    _lib.plan_chunks(0, 1, 1, 16)
  with pytest.raises(ValueError):
    _lib.plan_chunks(1, 1, 1, 16, lanes=5)


def test_numa_helper_parses_cpulists_and_degrades_gracefully():
  """se3ds_b200.hostmem: sysfs cpulist syntax; without a CUDA device nothing is bound."""
  from se3ds_b200 import hostmem
  assert hostmem._parse_cpulist('0-3,8,10-11') == {0, 1, 2, 3, 8, 10, 11}
  assert hostmem._parse_cpulist('') == set()
  assert hostmem._read('/nonexistent/se3ds') is None


def test_host_reprojector_rotates_workspaces_and_output_sets(monkeypatch):
  """guidance.HostReprojector without a device: the call order, which workspace / output set each batch gets,
  when the oldest call is waited for, and that a returned result is not the set the next submit writes."""
  from se3ds_b200 import guidance, _lib
  log = []

  class FakeWs:
    count = 0

    def __init__(self, device=0):
      FakeWs.count += 1
      self.name = 'ws%d' % FakeWs.count
      self.closed = False

    def host_wait(self):
      log.append(('wait', self.name))

    def close(self):
      self.closed = True

  def fake_reproject_host(rgb, depth, src_pos, tgt_pos, out=None, workspace=None, wait=True, **kw):
    assert wait is False and kw == {'device': 0, 'mask_frames': 1}
    out['batch'] = rgb
    log.append(('submit', workspace.name, id(out), rgb))
    return out

  monkeypatch.setattr(_lib, 'Workspace', FakeWs)
  monkeypatch.setattr(guidance, 'reproject_host', fake_reproject_host)
  pipe = guidance.HostReprojector(device=0, depth=2, mask_frames=1)
  got = []
  for b in range(5):
    done = pipe.submit(b, None, None, None)
    if b < 2:
      assert done is None  # the pipeline fills
    else:
      assert done['batch'] == b - 2  # oldest first
      newest = [e for e in log if e[0] == 'submit'][-1]
      assert newest[2] != id(done)  # the result just returned is not the set being written now
      got.append(done['batch'])
  got += [d['batch'] for d in pipe.flush()]
  assert got == [0, 1, 2, 3, 4]
  submits = [e for e in log if e[0] == 'submit']
  assert [e[1] for e in submits] == ['ws1', 'ws2', 'ws1', 'ws2', 'ws1']
  assert len({e[2] for e in submits}) == 3  # depth + 1 output sets
  # a workspace is waited for exactly before it is reused, and in submission order at the end
  assert [e[1] for e in log if e[0] == 'wait'] == ['ws1', 'ws2', 'ws1', 'ws2', 'ws1']
  assert log.index(('wait', 'ws1')) < log.index(submits[2])
  workspaces = list(pipe._ws)
  pipe.close()
  assert all(w.closed for w in workspaces)
  import pytest
  with pytest.raises(ValueError):
    guidance.HostReprojector(depth=0)
