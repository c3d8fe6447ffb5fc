"""GPU parity tests: the CUDA path (through the C ABI) against the canonical oracle.

Bar (BASELINE.json north_star): bit-exact target-pixel indices / z-buffer winners / masks, and
RGB / depth within 1e-5 relative.  The canonical arithmetic makes everything bit-exact, so the
assertions below use exact equality; the 1e-5 tolerance is only needed against the libm
restatement (tests/test_oracle_exact.py).
"""
import numpy as np
import pytest
import torch

from oracle import ref_exact as X
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu

F32 = np.float32


@pytest.fixture(scope='module')
def mods():
  from se3ds_b200 import _lib, guidance, synth
  from se3ds_b200.inference import perturbation_utils
  from se3ds_b200.utils import pano_utils, point_cloud_utils
  assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
  _lib.load()
  return dict(lib=_lib, g=guidance, synth=synth, pano=pano_utils, pc=point_cloud_utils, pert=perturbation_utils)


def _cuda(d):
  return {k: torch.as_tensor(v).cuda() for k, v in d.items()}


def _check_fused(mods, inp, conv=None, mask_frames=0, per_job_bin=False, ws=None):
  g = mods['g']
  conv = conv or g.EVAL_METRIC
  t = _cuda(inp)
  out = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=mask_frames,
                    unproject_void=conv.unproject_void, project_void=conv.project_void,
                    filter_void=conv.filter_void, per_job_bin=per_job_bin, return_winner=True, workspace=ws)
  torch.cuda.synchronize()
  if conv.filter_void:
    ref = _oracle_filtered(inp, mask_frames, per_job_bin)
  else:
    ref = _oracle_masked(inp, conv, mask_frames, per_job_bin)
  np.testing.assert_array_equal(out['winner'].cpu().numpy(), ref['winner'])
  np.testing.assert_array_equal(out['proj_depth'].cpu().numpy(), ref['depth'])
  np.testing.assert_array_equal(out['proj_mask'].cpu().numpy(), ref['mask'])
  np.testing.assert_array_equal(out['proj_image'].cpu().numpy(), ref['image'])
  # without winner indices the z-buffer key is 32-bit (depth only): the guidance must not change
  for key64 in (False, True):
    o2 = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=mask_frames,
                     unproject_void=conv.unproject_void, project_void=conv.project_void,
                     filter_void=conv.filter_void, per_job_bin=per_job_bin, workspace=ws, key64=key64)
    assert 'winner' not in o2
    for k in ('proj_depth', 'proj_mask', 'proj_image'):
      assert torch.equal(o2[k], out[k]), (k, key64)
  return out, ref


def _oracle_masked(inp, conv, mask_frames, per_job_bin):
  """X.reproject masks frame 0 only; emulate a prefix of masked frames by pre-masking."""
  rgb = inp['rgb'].astype(np.int32).copy()
  for k in range(mask_frames):
    rgb[:, k] = R.mask_pano(rgb[:, k], masked_region_value=-1)
  return X.reproject(rgb, inp['depth'], inp['src_pos'], inp['tgt_pos'], unproject_void=conv.unproject_void,
                     project_void=conv.project_void, mask_first_frame=False, per_job_bin=per_job_bin)


def _oracle_filtered(inp, mask_frames, per_job_bin):
  """SE3DSModel convention: compaction (models/models.py:229-237) before projection.  Winner
  indices are reported in the un-compacted numbering s*H*W + pixel, like the kernels do."""
  rgb = inp['rgb'].astype(np.int32).copy()
  n, s, h, w, _ = rgb.shape
  assert n == 1, 'the compaction reduces over the batch axis; SE3DSModel is batch 1'
  coords, feats, index = [], [], []
  for k in range(s):
    frame = rgb[:, k]
    if k < mask_frames:
      frame = R.mask_pano(frame, masked_region_value=-1)
    xyz1, f = X.equirectangular_to_pointcloud(frame, inp['depth'][:, k], -1, 20.0, interpolation_method='bilinear')
    xyz1 = xyz1 + np.concatenate([inp['src_pos'][:, k], np.zeros((n, 1), F32)], 1)[:, :, None]
    valid = np.any(f != -1, axis=(0, 2))
    coords.append(xyz1[:, :, valid]); feats.append(f[:, valid]); index.append(k * h * w + np.nonzero(valid)[0])
  coords = np.concatenate(coords, 2); feats = np.concatenate(feats, 1); index = np.concatenate(index)
  tgt = inp['tgt_pos'].reshape(n, -1, 3)
  outs = []
  for p in range(tgt.shape[1]):
    rel = coords - np.concatenate([tgt[:, p], np.zeros((n, 1), F32)], 1)[:, :, None]
    o = X.splat(rel, feats, h, w, 20.0, -1.0)
    lut = index if index.size else np.zeros(1, np.int64)
    o['winner'] = np.where(o['winner'] >= 0, lut[np.maximum(o['winner'], 0)], -1).astype(np.int32)
    outs.append(o)
  assert per_job_bin or len(outs) == 1
  o = {k: np.concatenate([x[k] for x in outs], 0) for k in ('depth', 'feat', 'winner')}
  image, d, mask = R.guidance_from_projection(o['depth'], o['feat'])
  return dict(image=image, depth=d, mask=mask, winner=o['winner'])


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dist', ['room', 'rand'])
@pytest.mark.parametrize('h', [64, 256])
def test_fused_single_frame(mods, dist, h):
  """Config c1 shape family: one pano, one pose (BASELINE.json configs[0])."""
  inp = mods['synth'].make_inputs(1, 1, 1, h, seed=h, dist=dist)
  out, ref = _check_fused(mods, inp)
  assert (ref['winner'] >= 0).mean() > 0.2


def test_fused_batch_global_bin(mods):
  """Whole-call reject bin: rejected points of every job land on job 0's pixel (0,0)."""
  inp = mods['synth'].make_inputs(3, 1, 1, 64, seed=1, dist='rand')
  out, ref = _check_fused(mods, inp)
  assert out['proj_image'][0, 0, 0].max().item() > 0.9


def test_fused_per_job_bin(mods):
  inp = mods['synth'].make_inputs(2, 1, 2, 64, seed=2, dist='rand')
  _check_fused(mods, inp, per_job_bin=True)


def test_fused_trajectory_and_mask(mods):
  """Config c3 shape family: S prior frames fused into one target; frame 0 masked
  (trainers/gan_manager.py:530-535)."""
  inp = mods['synth'].make_inputs(2, 4, 1, 64, seed=3, dist='room')
  for conv in (mods['g'].EVAL_METRIC, mods['g'].GAN_MANAGER):
    _check_fused(mods, inp, conv=conv, mask_frames=1)


def test_fused_pose_sweep(mods):
  """Config c4 shape family: P perturbed poses of one pano."""
  inp = mods['synth'].make_inputs(1, 1, 8, 64, seed=4, dist='room', sweep=True)
  _check_fused(mods, inp, per_job_bin=True)
  _check_fused(mods, inp, per_job_bin=False)


def test_fused_se3ds_model_convention(mods):
  """models/models.py flow: mask, compaction, void -1 on both sides, batch 1."""
  inp = mods['synth'].make_inputs(1, 3, 1, 64, seed=5, dist='room')
  _check_fused(mods, inp, conv=mods['g'].SE3DS_MODEL, mask_frames=3, per_job_bin=True)


def test_fused_int32_rgb_with_void_values(mods):
  """int32 RGB in [-1,255] as the rollout loops produce (gan_manager.py:541-543)."""
  inp = mods['synth'].make_inputs(2, 2, 1, 32, seed=6, dist='rand', rgb_dtype=np.int32)
  rng = np.random.default_rng(0)
  inp['rgb'][rng.uniform(size=inp['rgb'].shape) < 0.05] = -1
  for conv in (mods['g'].EVAL_METRIC, mods['g'].GAN_MANAGER):
    _check_fused(mods, inp, conv=conv, mask_frames=1)


@pytest.mark.parametrize('h', [3, 4, 5, 10])
def test_fused_tiny_and_ragged_sizes(mods, h):
  """W % 4 != 0 takes the scalar-load path; tiny panos exercise partially filled warps."""
  inp = mods['synth'].make_inputs(2, 2, 3, h, seed=h, dist='rand')
  _check_fused(mods, inp, mask_frames=1)


def test_fused_chunked_equals_unchunked(mods):
  """Job chunking (bounded workspace) must not change a bit, including the global bin."""
  inp = mods['synth'].make_inputs(3, 2, 3, 32, seed=7, dist='rand')
  ws = mods['lib'].Workspace(0, 0, 32 * 64 * (16 + 16) * 2)  # two jobs per chunk
  _check_fused(mods, inp, ws=ws)
  _check_fused(mods, inp, per_job_bin=True, ws=ws)
  ws1 = mods['lib'].Workspace(0, 0, 1)  # one job per chunk
  _check_fused(mods, inp, ws=ws1)
  ws.close(); ws1.close()


@pytest.mark.parametrize('lanes', [1, 2, 3, 4])
def test_fused_concurrent_lanes(mods, lanes):
  """Chunks dealt to concurrent streams (se3ds_ws_lanes) must not change a bit: global bin parked by
  every chunk and patched after the join, per-job bins, winner indices (64-bit keys), repeated calls
  on the same workspace, and a chunk count that is not a multiple of the lane count."""
  inp = mods['synth'].make_inputs(5, 2, 3, 32, seed=11, dist='rand')
  ws = mods['lib'].Workspace(0, 0, 32 * 64 * (16 + 16) * 2 * lanes)  # two jobs per chunk and lane
  ws.lanes(lanes, 1, 1)
  for _ in range(2):
    _check_fused(mods, inp, mask_frames=1, ws=ws)
  _check_fused(mods, inp, per_job_bin=True, ws=ws)
  room = mods['synth'].make_inputs(4, 1, 1, 64, seed=12, dist='room')
  _check_fused(mods, room, mask_frames=1, ws=ws)
  ws.close()


def test_fused_winner_at_owner_pixel(mods):
  """A rejected point nearer than every valid point of the reject bin's owner pixel takes that pixel's
  depth: no valid point is its winner then (-1).  Found by tests/tools/gpu_fuzz.py (gan_manager voids, a
  masked frame: many rejected points with small depths).  Single chunk, per-job bins, and the
  multi-chunk path where the owner pixel is patched after the last chunk."""
  g = mods['g']
  hit = 0
  for seed in range(6):
    inp = mods['synth'].make_inputs(2, 3, 1, 16, seed=100 + seed, dist='rand')
    for per_job in (False, True):
      _, ref = _check_fused(mods, inp, conv=g.GAN_MANAGER, mask_frames=1, per_job_bin=per_job)
      hit += int((ref['winner'][:, 0, 0] < 0).any() and (ref['depth'][:, 0, 0, 0] < 1).any())
    ws = mods['lib'].Workspace(0, 0, 1)  # one job per chunk: global bin parked, owner pixel patched at the end
    ws.lanes(2, 1, 1)
    _check_fused(mods, inp, conv=g.GAN_MANAGER, mask_frames=1, ws=ws)
    ws.close()
  assert hit > 0, 'no case exercised the owner-pixel rule'


def test_fused_lanes_argument_checks(mods):
  ws = mods['lib'].Workspace(0, 0, 0)
  for bad in (0, 5, -1):
    with pytest.raises(ValueError):
      ws.lanes(bad)
  with pytest.raises(ValueError):
    ws.lanes(2, -1)
  with pytest.raises(ValueError):
    ws.lanes(2, 0, -1)
  ws.close()


def _rotations(n, p, seed):
  rng = np.random.default_rng(seed)
  a = rng.standard_normal((n * p, 3, 3))
  qm, _ = np.linalg.qr(a)
  qm *= np.sign(np.linalg.det(qm))[:, None, None]
  return qm.reshape(n, p, 3, 3).astype(F32)


def test_se3_rotation(mods):
  """full SE(3) target poses inside the fused kernel.  R = I is the reference path
  bit for bit; random rotations match the oracle's canonical fma chain bit for bit and the float64
  rotation to within the usual boundary-case fraction."""
  g = mods['g']
  inp = mods['synth'].make_inputs(2, 2, 2, 128, seed=21, dist='room', sweep=True)
  t = _cuda(inp)
  eye = np.broadcast_to(np.eye(3, dtype=F32), (2, 2, 3, 3)).copy()
  base = {k: v.clone() for k, v in g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, return_winner=True).items()}
  same = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, return_winner=True, tgt_rot=eye)
  for k in base:
    assert torch.equal(base[k], same[k]), k
  rot = _rotations(2, 2, 3)
  out = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, return_winner=True, tgt_rot=rot)
  ref = X.reproject(inp['rgb'].astype(np.int32), inp['depth'], inp['src_pos'], inp['tgt_pos'], mask_first_frame=True, tgt_rot=rot)
  np.testing.assert_array_equal(out['winner'].cpu().numpy(), ref['winner'])
  np.testing.assert_array_equal(out['proj_depth'].cpu().numpy(), ref['depth'])
  np.testing.assert_array_equal(out['proj_mask'].cpu().numpy(), ref['mask'])
  np.testing.assert_array_equal(out['proj_image'].cpu().numpy(), ref['image'])
  # a rotation moves pixels but keeps every depth: the multiset of hit depths is (almost) unchanged
  d0 = np.sort(base['proj_depth'].cpu().numpy()[base['winner'].cpu().numpy() >= 0])
  d1 = np.sort(ref['depth'][ref['winner'] >= 0])
  assert abs(len(d0) - len(d1)) < 0.2 * len(d0)


def test_fused_identity_reprojection(mods):
  """models/models_test.py:64-68: project at the source position => >= 95 % RGB equal."""
  inp = mods['synth'].make_inputs(1, 1, 1, 128, seed=8, dist='rand')
  inp['tgt_pos'] = inp['src_pos'][:, 0].copy()
  out, _ = _check_fused(mods, inp, conv=mods['g'].SE3DS_MODEL, per_job_bin=True)
  proj_u8 = (out['proj_image'] * 255).to(torch.uint8).cpu().numpy()
  assert np.all(proj_u8[0] == inp['rgb'][0, 0], axis=-1).mean() >= 0.95


def test_fused_all_invalid_depth(mods):
  """Nothing valid: every output pixel is void (depth 1, mask 0, black)."""
  inp = mods['synth'].make_inputs(1, 1, 1, 16, seed=9)
  inp['depth'][:] = 0
  out, _ = _check_fused(mods, inp, conv=mods['g'].SE3DS_MODEL, per_job_bin=True)
  assert out['proj_mask'].sum().item() == 0
  assert (out['proj_depth'] == 1).all()


@pytest.mark.parametrize('h,dist', [(64, 'rand'), (256, 'room'), (512, 'room'), (512, 'rand'), (1024, 'room')])
def test_certified_fast_projection(mods, h, dist):
  """The default projection takes MUFU shortcuts only where their result is certified to equal
  the canonical one.  Verify mode evaluates both for every point: no certified point may differ,
  and the observed deviation must stay well inside the certification margin."""
  g, lib = mods['g'], mods['lib']
  n = 2 if h < 1024 else 1
  t = _cuda(mods['synth'].make_inputs(n, 2, 2, h, seed=h, dist=dist, sweep=True))
  ws = lib.Workspace(0)
  outs = {}
  for mode in (0, 1, 2):
    ws.projection_mode(mode)
    o = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, return_winner=True, workspace=ws)
    outs[mode] = {k: v.clone() for k, v in o.items()}
  v = ws.verify_read()
  ws.close()
  total = n * 2 * 2 * h * 2 * h  # void-feature points (masked rows, invalid depth) are never projected
  assert 0.5 * total < v['points'] <= total
  assert v['wrong'] == 0, v
  assert v['certified'] > 0.5 * v['points'], v
  # margins: dx = W * 1e-6, dy = 2 * H * 1e-6 pixels
  assert v['max_dev_x'] < 0.5 * (2 * h) * 1e-6 and v['max_dev_y'] < 0.5 * 2 * h * 1e-6, v
  for mode in (1, 2):
    for k in outs[0]:
      assert torch.equal(outs[0][k], outs[mode][k]), (mode, k)


@pytest.mark.parametrize('h', [500, 512, 2048])
def test_planted_column_borders(mods, h):
  """Adversarial input for the column certificate: source and target position coincide
  and the target frame is rotated about the vertical axis by (0.5 +- dx (1 + eps)) columns, so EVERY point
  sits at a column border +- dx (1 + eps) -- right at the certification margin, for all depths.  Verify
  mode: no certified point may differ from the canonical pixel; and the fast, canonical-only and verify
  results are identical."""
  g, lib = mods['g'], mods['lib']
  w = 2 * h
  inp = mods['synth'].make_inputs(1, 1, 1, h, seed=h + 3, dist='rand')
  inp['tgt_pos'] = inp['src_pos'][:, 0].copy()
  t = _cuda(inp)
  ws = lib.Workspace(0)
  dx = w * 1e-6
  for eps in (-0.3, -0.02, 0.02, 0.3):
    for sgn in (-1.0, 1.0):
      alpha = 2.0 * np.pi * (0.5 + sgn * dx * (1.0 + eps)) / w
      c, s_ = np.cos(alpha), np.sin(alpha)
      rot = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]], np.float64).astype(F32)[None, None]
      outs = {}
      for mode in (0, 1, 2):
        ws.projection_mode(mode)
        o = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], return_winner=True, workspace=ws, tgt_rot=rot)
        outs[mode] = {k: v.clone() for k, v in o.items()}
      v = ws.verify_read()
      assert v['wrong'] == 0, (eps, sgn, v)
      assert v['max_dev_x'] < 0.53 * dx, (eps, sgn, v)
      # the points really sit where they were planted: outside the margin they are certified, inside not
      if eps >= 0.3:
        assert v['certified'] > 0.9 * v['points'], (eps, sgn, v)
      if eps <= -0.3:
        assert v['certified'] < 0.1 * v['points'], (eps, sgn, v)
      for mode in (1, 2):
        for k in outs[0]:
          assert torch.equal(outs[0][k], outs[mode][k]), (mode, k, eps, sgn)
  ws.close()


def test_non_finite_inputs_do_not_corrupt_memory(mods):
  """NaN / Inf depth and positions are outside the contract (the reference propagates NaN through
  `(depth * scale) * mask`), but they must never index out of bounds or poison other items."""
  g = mods['g']
  inp = mods['synth'].make_inputs(2, 1, 1, 64, seed=20, dist='room')
  clean = g.reproject(*(torch.as_tensor(inp[k]).cuda() for k in ('rgb', 'depth', 'src_pos', 'tgt_pos')), per_job_bin=True)
  clean = {k: v.clone() for k, v in clean.items()}
  bad = {k: v.copy() for k, v in inp.items()}
  bad['depth'][0, 0, 10:20, 5:50] = np.nan
  bad['depth'][0, 0, 30:33, :] = np.inf
  bad['depth'][0, 0, 40, 7] = -np.inf
  out = g.reproject(*(torch.as_tensor(bad[k]).cuda() for k in ('rgb', 'depth', 'src_pos', 'tgt_pos')), per_job_bin=True,
                    return_winner=True)
  torch.cuda.synchronize()
  for k in ('proj_image', 'proj_depth', 'proj_mask'):
    assert torch.equal(out[k][1], clean[k][1]), k  # the untouched item is bit-identical
  w = out['winner'][0]
  assert int(w.max()) < 64 * 128 and int(w.min()) >= -1
  nanpos = {k: v.copy() for k, v in inp.items()}
  nanpos['tgt_pos'][0, 0, 1] = np.nan
  out = g.reproject(*(torch.as_tensor(nanpos[k]).cuda() for k in ('rgb', 'depth', 'src_pos', 'tgt_pos')), per_job_bin=True)
  torch.cuda.synchronize()
  assert out['proj_mask'][0, 1:].sum().item() == 0  # nothing projects from a NaN pose (pixel (0,0) holds the bin)
  for k in ('proj_image', 'proj_depth', 'proj_mask'):
    assert torch.equal(out[k][1], clean[k][1]), k


def test_export_and_apply_bin(mods):
  """Multi-GPU bin protocol on one GPU: export per shard, reduce, apply == whole-call result."""
  g = mods['g']
  inp = mods['synth'].make_inputs(4, 1, 1, 32, seed=10, dist='rand')
  ref = X.reproject(inp['rgb'], inp['depth'], inp['src_pos'], inp['tgt_pos'], mask_first_frame=False)
  t = _cuda(inp)
  parts, bins = [], []
  for lo, hi in ((0, 2), (2, 4)):
    o = g.reproject(t['rgb'][lo:hi], t['depth'][lo:hi], t['src_pos'][lo:hi], t['tgt_pos'][lo:hi], export_bin=True,
                    return_winner=True)
    bins.append(o['bin'].clone()); parts.append({k: v.clone() for k, v in o.items()})
  b = torch.stack(bins)
  red = torch.cat([b[:, :1].min(dim=0).values, b[:, 1:4].max(dim=0).values, bins[0][4:5]])
  g.apply_bin(red, parts[0])
  for k, r in (('proj_image', 'image'), ('proj_depth', 'depth'), ('proj_mask', 'mask'), ('winner', 'winner')):
    got = torch.cat([parts[0][k], parts[1][k]]).cpu().numpy()
    np.testing.assert_array_equal(got, ref[r])


def test_determinism(mods):
  """Atomic min / max are order independent: repeated runs are bit-stable."""
  g = mods['g']
  t = _cuda(mods['synth'].make_inputs(2, 2, 2, 128, seed=11, dist='rand'))
  first = {k: v.clone() for k, v in g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], return_winner=True).items()}
  for _ in range(3):
    again = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], return_winner=True)
    for k in first:
      assert torch.equal(first[k], again[k]), k


def test_prepared_plans_back_to_back(mods):
  """PreparedReprojection with SE3DS_FLAG_INPUTS_READY: consecutive runs overlap through programmatic
  dependent launch (K2's math starts before the previous resolve has finished) and must still give
  the results of isolated calls, also when two plans share one workspace."""
  g, lib = mods['g'], mods['lib']
  ws = lib.Workspace(0)
  plans, refs = [], []
  for seed in (31, 32, 33):
    t = _cuda(mods['synth'].make_inputs(2, 2, 1, 128, seed=seed, dist='rand'))
    ref = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, return_winner=True)
    refs.append({k: v.clone() for k, v in ref.items()})
    plans.append(g.prepare(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, return_winner=True,
                           workspace=ws, inputs_ready=True))
  for _ in range(5):
    for pl in plans:
      pl.run()
  torch.cuda.synchronize()
  for pl, ref in zip(plans, refs):
    for k in ref:
      assert torch.equal(pl.out[k], ref[k]), k
  ws.close()


@pytest.mark.parametrize('n,s,p,h', [(2, 2, 1, 64), (10, 1, 2, 32), (5, 2, 1, 16)])
def test_reproject_host_equals_device(mods, n, s, p, h):
  """The blocking host call pipelines groups of batch items (four stages): 2 items -> one per stage, 10 items ->
  groups of 3, 3, 3, 1 with the owner pixel patched after the last group; same bits as the device call."""
  g = mods['g']
  inp = mods['synth'].make_inputs(n, s, p, h, seed=12 + n, dist='rand', sweep=p > 1)
  t = _cuda(inp)
  dev = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, return_winner=True)
  host = g.reproject_host(*(torch.as_tensor(inp[k]) for k in ('rgb', 'depth', 'src_pos', 'tgt_pos')),
                          mask_frames=1, return_winner=True)
  for k in ('proj_image', 'proj_depth', 'proj_mask', 'winner'):
    assert not host[k].is_cuda
    assert torch.equal(host[k], dev[k].cpu()), k
  small = g.reproject_host(*(torch.as_tensor(inp[k]) for k in ('rgb', 'depth', 'src_pos', 'tgt_pos')), mask_frames=1, compact=True)
  back = g.expand_guidance({k: v.cuda() for k, v in small.items()})
  for k in ('proj_image', 'proj_depth', 'proj_mask'):
    assert torch.equal(back[k], dev[k]), k


@pytest.mark.parametrize('compact,depth', [(False, 2), (True, 2), (True, 3), (False, 1)])
def test_host_reprojector_pipeline_equals_blocking_calls(mods, compact, depth):
  """SE3DS_FLAG_HOST_ASYNC + se3ds_ws_host_wait behind guidance.HostReprojector: batches of different shapes
  and contents in flight on `depth` workspaces give, in submission order, exactly what the blocking call gives."""
  g = mods['g']
  shapes = [(2, 1, 1, 64), (3, 2, 2, 32), (1, 1, 3, 48), (2, 1, 1, 64), (4, 1, 1, 32), (2, 2, 1, 64), (1, 1, 1, 16)]
  batches = [mods['synth'].make_inputs(*sh, seed=70 + i, dist='rand', sweep=sh[2] > 1) for i, sh in enumerate(shapes)]
  host = [tuple(torch.as_tensor(b[k]).pin_memory() for k in ('rgb', 'depth', 'src_pos', 'tgt_pos')) for b in batches]
  want = [{k: v.clone() for k, v in g.reproject_host(*h, mask_frames=1, compact=compact).items()} for h in host]
  pipe = g.HostReprojector(device=0, depth=depth, mask_frames=1, compact=compact)
  got = []
  for h in host:
    done = pipe.submit(*h)
    if done is not None:
      got.append({k: v.clone() for k, v in done.items()})
  got += [{k: v.clone() for k, v in d.items()} for d in pipe.flush()]
  assert len(got) == len(want)
  for a, b in zip(got, want):
    assert a.keys() == b.keys()
    for k in b:
      assert torch.equal(a[k], b[k]), k
  # a blocking call on a workspace with a pending asynchronous call waits for it first
  ws = mods['lib'].Workspace(0)
  o1, o2 = {}, {}
  g.reproject_host(*host[0], mask_frames=1, compact=compact, workspace=ws, out=o1, wait=False)
  g.reproject_host(*host[1], mask_frames=1, compact=compact, workspace=ws, out=o2)
  ws.host_wait()  # nothing pending any more: a no-op
  for o, b in ((o1, want[0]), (o2, want[1])):
    for k in b:
      assert torch.equal(o[k], b[k]), k
  with pytest.raises(ValueError):
    g.reproject_host(*host[0], mask_frames=1, wait=False)
  ws.close()


def test_full_size_c2_against_oracle(mods):
  """BASELINE.json configs[1]: 512x1024, batch 8 -- full size, still bit-exact."""
  inp = mods['synth'].make_inputs(8, 1, 1, 512, seed=13, dist='room')
  out, ref = _check_fused(mods, inp, mask_frames=1)
  hit = (ref['winner'] >= 0).mean()
  assert 0.2 < hit < 1.0


def test_full_size_c3_c4_shapes_against_oracle(mods):
  """BASELINE.json configs[2] / configs[3] at full resolution, reduced batch: 4 prior frames fused
  into one target (gan_manager conventions), and a 16-pose sweep of one pano."""
  inp = mods['synth'].make_inputs(3, 4, 1, 512, seed=17, dist='room')
  _check_fused(mods, inp, conv=mods['g'].GAN_MANAGER, mask_frames=1)
  inp = mods['synth'].make_inputs(1, 1, 16, 512, seed=18, dist='room', sweep=True)
  _check_fused(mods, inp, per_job_bin=True)


def test_stress_resolution_2048x4096(mods):
  """BASELINE.json configs[4] resolution (8.4 M pixels per pano, 23-bit pixel indices, one job per
  chunk), two accumulated frames, against the oracle."""
  inp = mods['synth'].make_inputs(1, 2, 1, 2048, seed=19, dist='room')
  out, ref = _check_fused(mods, inp, mask_frames=1)
  assert (ref['winner'] >= 0).mean() > 0.3


# ------------------------------------------------------------------------------------------
# compat path: the reference's own function signatures
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('batch_size,image_size,multi', [(2, 64, False), (1, 128, False), (2, 64, True), (1, 128, True)])
def test_equirectangular_to_pointcloud(mods, batch_size, image_size, multi):
  """utils/pano_utils_test.py:89-111 + exact parity of xyz1."""
  rng = np.random.default_rng(1)
  shape = (batch_size, image_size, 2 * image_size) + ((3,) if multi else ())
  feats = rng.integers(0, R.NUM_MP3D_CLASSES, shape).astype(np.int32)
  depth = rng.uniform(0, R.DEPTH_SCALE, (batch_size, image_size, 2 * image_size)).astype(F32)
  depth[0, :4] = rng.uniform(0, 1, depth[0, :4].shape)
  xyz1, ff = mods['pano'].equirectangular_to_pointcloud(torch.as_tensor(feats), torch.as_tensor(depth), 0, 20.0)
  ex, ef = X.equirectangular_to_pointcloud(feats, depth, 0, 20.0)
  assert tuple(xyz1.shape) == (batch_size, 4, 2 * image_size**2)
  assert ff.dtype == torch.int32
  np.testing.assert_array_equal(xyz1.cpu().numpy(), ex)
  np.testing.assert_array_equal(ff.cpu().numpy(), ef)
  _, fb = mods['pano'].equirectangular_to_pointcloud(torch.as_tensor(feats), torch.as_tensor(depth), 0, 20.0,
                                                     interpolation_method='bilinear')
  assert fb.dtype == torch.float32


@pytest.mark.parametrize('size_mult,method', [(2.0, 'nearest'), (1.5, 'bilinear'), (0.5, 'nearest'), (0.75, 'bilinear')])
def test_equirectangular_to_pointcloud_size_mult(mods, size_mult, method):
  """utils/pano_utils.py:203-208: denser / sparser clouds through tf.image.resize semantics
  (half-pixel centres; depth 'nearest', features with the given method)."""
  rng = np.random.default_rng(8)
  feats = rng.integers(0, 256, (2, 16, 32, 3)).astype(np.int32)
  depth = rng.uniform(0, 1.05, (2, 16, 32)).astype(F32)
  xyz1, ff = mods['pano'].equirectangular_to_pointcloud(torch.as_tensor(feats), torch.as_tensor(depth), -1, 20.0,
                                                        size_mult=size_mult, interpolation_method=method)
  ex, ef = X.equirectangular_to_pointcloud(feats, depth, -1, 20.0, size_mult=size_mult, interpolation_method=method)
  m = int(16 * size_mult) * int(32 * size_mult)
  assert tuple(xyz1.shape) == (2, 4, m) and tuple(ff.shape) == (2, m, 3)
  assert ff.dtype == (torch.int32 if method == 'nearest' else torch.float32)
  np.testing.assert_array_equal(xyz1.cpu().numpy(), ex)
  np.testing.assert_array_equal(ff.cpu().numpy(), ef)


@pytest.mark.parametrize('batch_size,image_size', [(2, 64), (1, 128)])
def test_project_feats_to_equirectangular(mods, batch_size, image_size):
  """utils/pano_utils_test.py:67-87 (random normal cloud, scalar semantic features)."""
  rng = np.random.default_rng(2)
  m = image_size**2
  feats = rng.integers(0, R.NUM_MP3D_CLASSES, (batch_size, m)).astype(np.int32)
  xyz = rng.standard_normal((batch_size, 3, m)).astype(F32)
  xyz1 = np.concatenate([xyz, np.ones((batch_size, 1, m), F32)], axis=1)
  d, f, win = mods['pano'].project_feats_to_equirectangular(torch.as_tensor(feats), torch.as_tensor(xyz1), image_size,
                                                            image_size * 2, 0, 20.0, return_winner=True)
  o = X.splat(xyz1, feats, image_size, image_size * 2, 20.0, 0.0)
  assert tuple(d.shape) == (batch_size, image_size, image_size * 2) and tuple(f.shape) == tuple(d.shape)
  np.testing.assert_array_equal(d.cpu().numpy(), o['depth'])
  np.testing.assert_array_equal(f.cpu().numpy(), o['feat'])
  np.testing.assert_array_equal(win.cpu().numpy(), o['winner'])
  assert 0 <= d.min().item() and d.max().item() <= 1
  assert 0 <= f.min().item() and f.max().item() <= R.NUM_MP3D_CLASSES


@pytest.mark.parametrize('batch_size,image_size,multi', [(2, 64, False), (1, 128, True)])
def test_project_to_feat_perspective(mods, batch_size, image_size, multi):
  """utils/point_cloud_utils_test.py:42-64: square target through project_to_feat itself."""
  rng = np.random.default_rng(3)
  shape = (batch_size, image_size, image_size) + ((3,) if multi else ())
  feats = rng.integers(0, R.NUM_MP3D_CLASSES, shape).astype(np.int32)
  depth = rng.uniform(0, R.DEPTH_SCALE, (batch_size, image_size, image_size)).astype(F32)
  xyz1, ff = R.get_filtered_coords_and_feats(feats, depth, 20.0)
  g_xyz1, g_ff = mods['pc'].get_filtered_coords_and_feats(torch.as_tensor(feats), torch.as_tensor(depth), 20.0)
  np.testing.assert_allclose(g_xyz1.cpu().numpy(), xyz1, rtol=1e-6, atol=1e-6)
  np.testing.assert_array_equal(g_ff.cpu().numpy(), ff)
  pd, pf = mods['pc'].project_to_feat(torch.as_tensor(xyz1), torch.as_tensor(ff), image_size, image_size, 20.0, 0)
  ed, ef = X.project_to_feat(xyz1, ff, image_size, image_size, 20.0, 0)
  np.testing.assert_array_equal(pd.cpu().numpy(), ed)
  np.testing.assert_array_equal(pf.cpu().numpy(), ef)
  assert tuple(pf.shape) == shape


def test_project_cloud_float_feats_negative_and_void_out(mods):
  """float32 features of mixed sign, non-zero output void class, C = 5."""
  rng = np.random.default_rng(4)
  n, m, c, h = 2, 5000, 5, 32
  xyz1 = np.concatenate([rng.standard_normal((n, 3, m)).astype(F32) * 3, np.ones((n, 1, m), F32)], 1)
  feats = rng.standard_normal((n, m, c)).astype(F32) * 10
  feats[rng.uniform(size=(n, m)) < 0.1] = -7.0
  coords = np.stack([xyz1[:, 0], xyz1[:, 1], np.abs(xyz1[:, 2]) + 0.1, xyz1[:, 3]], 1)
  # transformed-coordinates mode with an explicit output void class
  pd, pf = mods['pc'].project_to_feat(torch.as_tensor(coords), torch.as_tensor(feats), h, h, 20.0, -7.0, -3.0)
  ed, ef = X.project_to_feat(coords, feats, h, h, 20.0, -7.0, -3.0)
  np.testing.assert_array_equal(pd.cpu().numpy(), ed)
  np.testing.assert_array_equal(pf.cpu().numpy(), ef)


def test_project_empty_cloud(mods):
  """M = 0: the first frame of a rollout projects an empty memory (gan_manager.py:462-483)."""
  d, f = mods['pano'].project_feats_to_equirectangular(torch.zeros((2, 0, 3), dtype=torch.int32),
                                                       torch.zeros((2, 4, 0)), 16, 32, -1, 20.0)
  assert (d == 1).all() and (f == 0).all() and tuple(f.shape) == (2, 16, 32, 3)


def test_compat_pipeline_matches_fused(mods):
  """The op-by-op reference call sequence (materialised cloud) and the fused call agree."""
  g, pano = mods['g'], mods['pano']
  inp = mods['synth'].make_inputs(2, 2, 1, 64, seed=14, dist='room')
  t = _cuda(inp)
  coords, feats = [], []
  for k in range(2):
    frame = t['rgb'][:, k].to(torch.int32)
    if k == 0:
      frame = pano.mask_pano(frame, masked_region_value=-1)
    xyz1, f = pano.equirectangular_to_pointcloud(frame, t['depth'][:, k], -1, 20.0)
    pos = torch.cat([t['src_pos'][:, k], torch.zeros_like(t['src_pos'][:, k, :1])], 1)
    coords.append(xyz1 + pos[:, :, None]); feats.append(f)
  coords = torch.cat(coords, 2); feats = torch.cat(feats, 1)
  tp = torch.cat([t['tgt_pos'][:, 0], torch.zeros_like(t['tgt_pos'][:, 0, :1])], 1)
  d, f = pano.project_feats_to_equirectangular(feats, coords - tp[:, :, None], 64, 128, -1, 20.0)
  out = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1)
  assert torch.equal(out['proj_depth'][..., 0], d)
  # numpy's float32 division is IEEE; torch's CUDA `f / 255` multiplies by a reciprocal
  np.testing.assert_array_equal(out['proj_image'].cpu().numpy(), np.clip(f.cpu().numpy() / F32(255), 0, 1))


@pytest.mark.parametrize('batch_size,h,dtype', [(2, 64, torch.float32), (2, 64, torch.int32), (1, 256, torch.int32), (1, 8, torch.uint8)])
def test_mask_pano(mods, batch_size, h, dtype):
  """utils/pano_utils_test.py:113-123 + exact parity."""
  rng = np.random.default_rng(5)
  pano = torch.as_tensor(rng.uniform(0, 255, (batch_size, h, 2 * h, 3))).to(dtype)
  for value in (0, -1 if dtype != torch.uint8 else 7):
    out = mods['pano'].mask_pano(pano, masked_region_value=value)
    assert out.shape == pano.shape and out.dtype == pano.dtype
    np.testing.assert_array_equal(out.cpu().numpy(), R.mask_pano(pano.numpy(), masked_region_value=value))


@pytest.mark.parametrize('distance,depth_distance,expected', [(0.5, 0.5, 1.0), (0.3, 0.5, 0.0)])
def test_proportion_invalid_kat(mods, distance, depth_distance, expected):
  """inference/perturbation_utils_test.py:30-41."""
  depth = torch.full((64, 128), depth_distance / 20.0)
  got = mods['pert'].get_proportion_invalid_for_depth(torch.tensor([0.0, distance, 0.0]), depth)
  assert got == expected


def test_proportion_invalid_offsets_and_batch(mods):
  """inference/perturbation_utils_test.py:43-94 + 64 random offsets against the oracle."""
  pert = mods['pert']
  for offset, centre in (([0.0, 0.5, 0.0], (0.5, 0.5)), ([0.5, 0.5, 0.0], (0.75, 0.75))):
    img = np.full((64, 128), 1.0, F32)
    hs, ws = int(64 * centre[0]), int(128 * centre[1])
    img[hs - 10:hs + 10, ws - 10:ws + 10] = 0.0
    assert pert.get_proportion_invalid_for_depth(torch.tensor(offset), torch.as_tensor(img)) > 0.0
    img = np.full((64, 128), 1.0, F32)
    img[:10, :10] = 0.0
    assert pert.get_proportion_invalid_for_depth(torch.tensor(offset), torch.as_tensor(img)) == 0.0
  rng = np.random.default_rng(6)
  depth = rng.uniform(0, 0.2, (128, 256)).astype(F32)
  offs = np.concatenate([rng.uniform(-1.5, 1.5, (64, 2)), rng.uniform(-0.1, 0.1, (64, 1))], 1).astype(F32)
  got = pert.get_proportion_invalid_for_depth_batch(torch.as_tensor(offs), torch.as_tensor(depth)).cpu().numpy()
  want = np.array([X.get_proportion_invalid_for_depth(o, depth) for o in offs], F32)
  np.testing.assert_array_equal(got, want)


def test_guidance_memory_matches_reference_flow(mods):
  """GuidanceMemory (frame ring) against the oracle's SE3DSModel memory flow."""
  g = mods['g']
  inp = mods['synth'].make_inputs(1, 2, 1, 64, seed=15, dist='room')
  rng = np.random.default_rng(7)
  sem = rng.integers(1, R.NUM_MP3D_CLASSES, (1, 2, 64, 128, 1)).astype(np.uint8)
  mem = g.GuidanceMemory(64, project_semantic=True)
  ora = R.SE3DSMemoryOracle(64)
  for k in range(2):
    mem.add_to_memory(torch.as_tensor(inp['rgb'][:, k]), torch.as_tensor(sem[:, k]), torch.as_tensor(inp['depth'][:, k]),
                      torch.as_tensor(inp['src_pos'][:, k]))
    ora.add_to_memory(inp['rgb'][:, k], sem[:, k], inp['depth'][:, k], inp['src_pos'][:, k])
  out = mem(torch.as_tensor(inp['tgt_pos'][:, 0]))
  want = ora.project(inp['tgt_pos'][:, 0])  # libm arithmetic: boundary cases may differ
  for k in ('proj_image', 'proj_depth', 'proj_mask'):
    got = out[k].cpu().numpy()
    # libm sin/cos differ from the canonical tables by an ulp, so depth agrees to 1e-5 relative
    # (north_star tolerance) rather than bit for bit; indices / colours flip only at pixel borders.
    bad = np.mean(np.any(np.abs(got - want[k]) > 1e-5 * np.abs(want[k]), axis=-1))
    assert bad < 1e-3, (k, bad)
  assert np.mean(out['proj_semantic'].cpu().numpy() != want['proj_semantic']) < 1e-3
  with pytest.raises(ValueError):
    g.GuidanceMemory(64, batch_size=2)


def test_fused_semantic_projection(mods):
  """models/models.py:217-219,229-231,276-278: scalar class ids through the fused kernels (ids
  replicated into three channels, void class 0, compaction, raw features) against the canonical
  oracle on the compacted cloud, bit for bit."""
  g = mods['g']
  inp = mods['synth'].make_inputs(1, 3, 1, 64, seed=23, dist='room')
  rng = np.random.default_rng(11)
  sem = rng.integers(0, R.NUM_MP3D_CLASSES, (1, 3, 64, 128)).astype(np.uint8)  # includes void (0) pixels
  mem = g.GuidanceMemory(64, project_semantic=True)
  coords, feats = [], []
  for k in range(3):
    mem.add_to_memory(torch.as_tensor(inp['rgb'][:, k]), torch.as_tensor(sem[:, k, ..., None]),
                      torch.as_tensor(inp['depth'][:, k]), torch.as_tensor(inp['src_pos'][:, k]))
    xyz1, f = X.equirectangular_to_pointcloud(sem[:, k], inp['depth'][:, k], 0, 20.0)
    xyz1 = xyz1 + np.concatenate([inp['src_pos'][:, k], np.zeros((1, 1), F32)], 1)[:, :, None]
    valid = np.any(f != 0, axis=0)
    coords.append(xyz1[:, :, valid]); feats.append(f[:, valid])
  out = mem(torch.as_tensor(inp['tgt_pos'][:, 0]))
  rel = np.concatenate(coords, 2) - np.concatenate([inp['tgt_pos'][:, 0], np.zeros((1, 1), F32)], 1)[:, :, None]
  want = X.splat(rel, np.concatenate(feats, 1), 64, 128, 20.0, 0.0)['feat']
  np.testing.assert_array_equal(out['proj_semantic'].cpu().numpy(), want.astype(np.uint8))
  assert out['proj_semantic'].dtype == torch.uint8 and (want > 0).mean() > 0.2


def test_ply_export(mods, tmp_path):
  """models/models.py:154-178: the only on-disk format on the path."""
  g = mods['g']
  inp = mods['synth'].make_inputs(1, 2, 1, 16, seed=16, dist='room')
  sem = np.ones((1, 16, 32, 1), np.uint8)
  mem = g.GuidanceMemory(16)
  ora = R.SE3DSMemoryOracle(16)
  for k in range(2):
    mem.add_to_memory(torch.as_tensor(inp['rgb'][:, k]), torch.as_tensor(sem), torch.as_tensor(inp['depth'][:, k]),
                      torch.as_tensor(inp['src_pos'][:, k]))
    ora.add_to_memory(inp['rgb'][:, k], sem, inp['depth'][:, k], inp['src_pos'][:, k])
  path = tmp_path / 'cloud.ply'
  mem.write_memory_as_pointcloud(str(path))
  lines = path.read_text().splitlines()
  state = ora.get_memory_state()
  m = state.rgb_coords.shape[2]
  assert lines[:3] == ['ply', 'format ascii 1.0 ', 'element vertex %d' % m] and lines[9] == 'end_header'
  assert len(lines) == 10 + m
  vals = np.array([l.split() for l in lines[10:]], dtype=np.float64)
  np.testing.assert_allclose(vals[:, :3], state.rgb_coords[0, :3].T, atol=1e-5)
  np.testing.assert_array_equal(vals[:, 3:].astype(np.int32), state.rgb[0])


def test_error_behaviour(mods):
  """Same exception types as the reference (pano_utils.py:190-202, point_cloud_utils.py:120-122)."""
  pano, pc = mods['pano'], mods['pc']
  with pytest.raises(ValueError):
    pano.equirectangular_to_pointcloud(torch.zeros((1, 4, 8, 3, 1), dtype=torch.int32), torch.zeros((1, 4, 8)), 0, 20.0)
  with pytest.raises(ValueError):
    pano.equirectangular_to_pointcloud(torch.zeros((1, 4, 8, 3), dtype=torch.uint8), torch.zeros((1, 4, 8)), -1, 20.0)
  with pytest.raises(AssertionError):
    pano.equirectangular_to_pointcloud(torch.zeros((1, 4, 9, 3), dtype=torch.int32), torch.zeros((1, 4, 9)), -1, 20.0)
  with pytest.raises(ValueError):
    pc.project_to_feat(torch.zeros((1, 4, 5)), torch.zeros((1, 5, 3, 1)), 8, 8, 20.0, 0)


def _quantize(image):
  """trainers/gan_manager.py:539-542: clip(int32(image * 255), -1, 255), truncating cast."""
  return np.clip(np.trunc(image.astype(F32) * F32(255)).astype(np.int64), -1, 255).astype(np.int32)


@pytest.mark.parametrize('conv_name,feed_depth', [('GAN_MANAGER', True), ('EVAL_METRIC', True), ('EVAL_METRIC', False)])
def test_rollout_matches_oracle_loop(mods, conv_name, feed_depth):
  """`guidance.rollout` owns the trajectory loop of trainers/gan_manager.py:458-556 /
  utils/eval_metric.py:144-240 (empty memory at frame 0, frame-0 mask, generated frames fed back clipped to
  [-1, 255], depth fed back).  The generator is a deterministic stand-in; every frame's guidance must equal
  the oracle's projection of the memory the reference loop would have built, bit for bit."""
  g = mods['g']
  conv = getattr(g, conv_name)
  n, t_all, h = 2, 4, 32
  w = 2 * h
  inp = mods['synth'].make_inputs(n, t_all, 1, h, seed=31, dist='room')
  images = (inp['rgb'].astype(F32) / F32(255)).astype(F32)           # (N,T,H,W,3) in [0,1]
  depths = inp['depth'][..., None]                                   # (N,T,H,W,1)
  positions = inp['src_pos']                                         # (N,T,3)

  def stand_in(proj_image, proj_mask, gt, gt_depth, t):
    """'Generator': keeps the projected pixels, fills the holes from the ground truth (slightly dimmed,
    with a few values pushed outside [0, 1] so that the clip of the feedback path matters)."""
    gen = np.where(proj_mask > 0, proj_image, gt * F32(0.9)).astype(F32)
    gen[:, ::7, ::5] = F32(1.2)
    gen[:, 3::11, 1::6] = F32(-0.3)
    dep = (gt_depth * F32(1.0 - 0.01 * t)).astype(F32)
    return gen, dep

  calls = []
  def generator_fn(inputs, t):
    calls.append({k: v.clone() for k, v in inputs.items()})
    gen, dep = stand_in(inputs['proj_image'].cpu().numpy(), inputs['proj_mask'].cpu().numpy(), images[:, t], depths[:, t], t)
    return torch.as_tensor(gen).cuda(), torch.as_tensor(dep).cuda()

  res = g.rollout(torch.as_tensor(images), torch.as_tensor(depths), torch.as_tensor(positions), generator_fn,
                  convention=conv, feed_depth=feed_depth)
  torch.cuda.synchronize()
  assert len(res['guidance']) == t_all and len(calls) == t_all

  # the oracle's loop: a frame list instead of the concatenated cloud (same points, same order)
  mem_rgb, mem_depth, mem_pos = [], [], []
  prev = np.zeros_like(images[:, 0])
  for t in range(t_all):
    if t == 0:
      want = dict(image=np.zeros((n, h, w, 3), F32), depth=np.ones((n, h, w, 1), F32), mask=np.zeros((n, h, w, 1), F32))
    else:
      want = X.reproject(np.stack(mem_rgb, 1), np.stack(mem_depth, 1), np.stack(mem_pos, 1), positions[:, t],
                         unproject_void=conv.unproject_void, project_void=conv.project_void, mask_first_frame=True)
    got = res['guidance'][t]
    np.testing.assert_array_equal(got['proj_image'].cpu().numpy(), want['image'], err_msg=f'frame {t}')
    np.testing.assert_array_equal(got['proj_depth'].cpu().numpy(), want['depth'], err_msg=f'frame {t}')
    np.testing.assert_array_equal(got['proj_mask'].cpu().numpy(), want['mask'], err_msg=f'frame {t}')
    assert (got['blurred_mask'] == 0).all()
    np.testing.assert_array_equal(calls[t]['prev_image'].cpu().numpy(), prev)
    assert calls[t]['first_frame'].cpu().numpy().tolist() == [1.0 if t == 0 else 0.0] * n
    gen, dep = stand_in(want['image'], want['mask'], images[:, t], depths[:, t], t)
    if t == 0:
      prev = images[:, 0]
      mem_rgb.append(_quantize(images[:, 0])); mem_depth.append(depths[:, 0, ..., 0])
    else:
      prev = gen
      mem_rgb.append(_quantize(gen)); mem_depth.append(dep[..., 0] if feed_depth else depths[:, t, ..., 0])
    mem_pos.append(positions[:, t])
  assert res['guidance'][2]['proj_mask'].mean().item() > 0.3  # the later frames really see the memory
  ring = res['memory']
  np.testing.assert_array_equal(ring.rgb[:, :ring.count].cpu().numpy(), np.stack(mem_rgb, 1))


def test_guidance_memory_pose_sweep_and_states(mods):
  """VERDICT r1 missing 2 + 3: GuidanceMemory renders P target positions in one call (the VLN sweep of
  inference/perturbation_utils.py + models/models.py:247-321), keeps its frames in a ring between calls,
  and converts to / from the reference's MemoryState (models/models.py:77-87)."""
  g = mods['g']
  h = 32
  inp = mods['synth'].make_inputs(1, 3, 6, h, seed=41, dist='room', sweep=True)
  rng = np.random.default_rng(3)
  sem = rng.integers(0, R.NUM_MP3D_CLASSES, (1, 3, h, 2 * h, 1)).astype(np.uint8)
  mem = g.GuidanceMemory(h, project_semantic=True)
  ora = R.SE3DSMemoryOracle(h)
  masked = [False, True, True]   # a masked frame after an unmasked one: the ring reorders, nothing else changes
  for k in range(3):
    mem.add_to_memory(torch.as_tensor(inp['rgb'][:, k]), torch.as_tensor(sem[:, k]), torch.as_tensor(inp['depth'][:, k]),
                      torch.as_tensor(inp['src_pos'][:, k]), mask_blurred=masked[k])
    ora.add_to_memory(inp['rgb'][:, k], sem[:, k], inp['depth'][:, k], inp['src_pos'][:, k], mask_blurred=masked[k])
  poses = torch.as_tensor(inp['tgt_pos'][0])                         # (P,3)
  sweep = {k: v.clone() for k, v in mem(poses).items()}
  assert sweep['proj_image'].shape == (6, h, 2 * h, 3)
  for p in range(6):                                                  # P poses at once == one call per pose
    one = mem(poses[p])
    for k in one:
      assert torch.equal(one[k][0], sweep[k][p]), (k, p)
  # reference-shaped state: the same point SETS as the oracle's model (the ring keeps masked frames in front,
  # so the order of the frames may differ; min / max do not care)
  st = mem.to_reference_state()
  ost = ora.get_memory_state()
  assert st.rgb_coords.shape == ost.rgb_coords.shape and st.coords.shape == ost.coords.shape
  def rows(coords, feats):
    a = np.concatenate([coords[0].T.astype(np.float64), feats[0].reshape(feats.shape[1], -1).astype(np.float64)], 1)
    return a[np.lexsort(a.T[::-1])]
  np.testing.assert_allclose(rows(st.rgb_coords.cpu().numpy(), st.rgb.cpu().numpy()), rows(ost.rgb_coords, ost.rgb), atol=1e-5)
  np.testing.assert_allclose(rows(st.coords.cpu().numpy(), st.feats.cpu().numpy()), rows(ost.coords, ost.feats), atol=1e-5)
  # cloud mode: adopt the reference state, project like the reference does -- identical guidance, bit for bit
  mem2 = g.GuidanceMemory(h, project_semantic=True)
  mem2.from_reference_state(st)
  cloud = mem2(poses)
  for k in ('proj_image', 'proj_depth', 'proj_mask', 'proj_semantic', 'blurred_mask'):
    assert torch.equal(cloud[k], sweep[k]), k
  # ... and it keeps growing like the reference memory
  extra = mods['synth'].make_inputs(1, 1, 1, h, seed=43, dist='room')
  for m_ in (mem, mem2):
    m_.add_to_memory(torch.as_tensor(extra['rgb'][:, 0]), torch.as_tensor(sem[:, 0]), torch.as_tensor(extra['depth'][:, 0]),
                     torch.as_tensor(extra['src_pos'][:, 0]), mask_blurred=False)
  a, b = mem(poses[:2]), mem2(poses[:2])
  for k in a:
    assert torch.equal(a[k], b[k]), k
  # native state round trip
  mem3 = g.GuidanceMemory(h, project_semantic=True)
  mem3.set_memory_state(mem.get_memory_state())
  c = mem3(poses[:2])
  for k in a:
    assert torch.equal(a[k], c[k]), k


@pytest.mark.parametrize('n,s,p,h,chunk_mb,per_job', [(2, 1, 1, 64, 0, False), (3, 2, 2, 32, 1, False), (2, 1, 3, 37, 0, True)])
def test_compact_output_expands_to_the_float32_contract(mods, n, s, p, h, chunk_mb, per_job):
  """SE3DS_FLAG_COMPACT_OUT (uint8 colours + float32 depth, 7 B instead of 20 B per pixel) followed by
  se3ds_expand_guidance is bit-identical to the float32 outputs: single chunk, several chunks with the
  owner-pixel patch, per-job bins, the host entry point, and the export / apply bin protocol."""
  g, lib = mods['g'], mods['lib']
  inp = mods['synth'].make_inputs(n, s, p, h, seed=51 + h, dist='rand', sweep=p > 1)
  t = _cuda(inp)
  ws = lib.Workspace(0, 0, chunk_mb << 20) if chunk_mb else lib.Workspace(0)
  kw = dict(mask_frames=1, per_job_bin=per_job, workspace=ws)
  want = {k: v.clone() for k, v in g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], **kw).items()}
  got = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], compact=True, **kw)
  assert got['proj_rgb_u8'].dtype == torch.uint8 and 'proj_mask' not in got
  got = g.expand_guidance({k: v.clone() for k, v in got.items()})
  for k in ('proj_image', 'proj_depth', 'proj_mask'):
    assert torch.equal(got[k], want[k]), k
  # host entry point: compact tensors come back over PCIe
  hout = g.reproject_host(inp['rgb'], inp['depth'], inp['src_pos'], inp['tgt_pos'], mask_frames=1, per_job_bin=per_job,
                          compact=True, workspace=ws)
  exp = g.expand_guidance({k: v.cuda() for k, v in hout.items()})
  for k in ('proj_image', 'proj_depth', 'proj_mask'):
    assert torch.equal(exp[k], want[k]), k
  # a job map re-orders the jobs on the way
  j = n * p
  perm = torch.randperm(j)
  shuffled = {'proj_rgb_u8': got['proj_rgb_u8'][perm].contiguous(), 'proj_depth': got['proj_depth'][perm].contiguous()}
  back = g.expand_guidance(shuffled, job_map=perm.to(torch.int32))   # slot s holds job perm[s]
  for k in ('proj_image', 'proj_depth', 'proj_mask'):
    assert torch.equal(back[k], want[k]), k
  if not per_job:  # export / apply bin with compact tensors
    a = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, export_bin=True, compact=True, workspace=ws)
    a = {k: v.clone() for k, v in a.items()}
    g.apply_bin(a.pop('bin'), a)
    a = g.expand_guidance(a)
    for k in ('proj_image', 'proj_depth', 'proj_mask'):
      assert torch.equal(a[k], want[k]), k
  ws.close()


def test_int32_colour_range_is_guarded(mods):
  """the fused path reduces colours in float16 (exact on [-2048, 2048]); values outside are
  refused instead of being rounded silently; inside the range int32 inputs are exact."""
  g = mods['g']
  inp = mods['synth'].make_inputs(1, 1, 1, 16, seed=3, dist='rand')
  rgb = inp['rgb'].astype(np.int32)
  rgb[0, 0, 5, 7] = 4097
  with pytest.raises(ValueError, match='2048'):
    g.reproject(torch.as_tensor(rgb), torch.as_tensor(inp['depth']), torch.as_tensor(inp['src_pos']), torch.as_tensor(inp['tgt_pos']))
  rgb[0, 0, 5, 7] = 2047
  rgb[0, 0, 6, 7] = -2000
  out = g.reproject(torch.as_tensor(rgb), torch.as_tensor(inp['depth']), torch.as_tensor(inp['src_pos']), torch.as_tensor(inp['tgt_pos']),
                    raw_features=True, return_winner=True)
  ref = X.reproject(rgb, inp['depth'], inp['src_pos'], inp['tgt_pos'], mask_first_frame=False)
  np.testing.assert_array_equal(out['proj_image'].cpu().numpy(), ref['raw_rgb'])
  np.testing.assert_array_equal(out['winner'].cpu().numpy(), ref['winner'])


@pytest.mark.parametrize('dtype,with_winner', [(np.int32, True), (np.int32, False), (np.uint8, False)])
def test_project_cloud_rgb_fast_accumulator(mods, dtype, with_winner):
  """The drop-in call of models/models.py:276-281 (3-channel integer colours, void -1 in, 0 out) takes the tuned compat
  kernels: float16x4 reduction, 32-bit key without winner indices, certified fast projection.  Bit-exact against the
  oracle, including values float16 cannot hold (they go to the float32 side accumulator) and a second call on the
  same workspace (everything re-armed)."""
  rng = np.random.default_rng(12)
  h, n = 48, 2
  inp = mods['synth'].make_inputs(n, 1, 1, h, seed=77, dist='rand')
  rgb = inp['rgb'][:, 0].astype(np.int32)
  if dtype == np.int32:
    rgb[:, ::5, ::7] = -1                      # void pixels
    rgb[0, 3::9, 2::11, 1] = 70000             # float16 cannot hold these
    rgb[1, 4::13, 1::5, 2] = -3000
  xyz1, feats = X.equirectangular_to_pointcloud(rgb, inp['depth'][:, 0], -1 if dtype == np.int32 else 0, 20.0)
  xyz1 = (xyz1 + np.concatenate([inp['src_pos'][:, 0] - inp['tgt_pos'][:, 0], np.zeros((n, 1), F32)], 1)[:, :, None]).astype(F32)
  feats = feats.astype(dtype)
  void_in = -1 if dtype == np.int32 else 0
  o = X.splat(xyz1, feats.astype(F32), h, 2 * h, 20.0, float(void_in))
  for _ in range(2):
    res = mods['pano'].project_feats_to_equirectangular(torch.as_tensor(feats), torch.as_tensor(xyz1), h, 2 * h, void_in, 20.0,
                                                        return_winner=with_winner)
    np.testing.assert_array_equal(res[0].cpu().numpy(), o['depth'])
    np.testing.assert_array_equal(res[1].cpu().numpy(), o['feat'])
    if with_winner:
      np.testing.assert_array_equal(res[2].cpu().numpy(), o['winner'])
  assert (o['feat'] > 60000).any() or dtype == np.uint8


@pytest.mark.parametrize('conv_name', ['EVAL_METRIC', 'GAN_MANAGER'])
def test_nan_depth_never_enters_the_reject_bin(mods, conv_name):
  """Regression (found by tests/tools/gpu_fuzz.py in round 2): a NaN depth makes a NaN radius; scatter-min ignores
  it, so it must not lower the reject bin's minimum depth either.  Full bit parity with 20 % NaN depths, masked
  and unmasked rows, both bin modes."""
  inp = mods['synth'].make_inputs(1, 1, 2, 64, seed=13, dist='room', sweep=True)
  rng = np.random.default_rng(4)
  inp['depth'][rng.random(inp['depth'].shape) < 0.2] = np.nan
  conv = getattr(mods['g'], conv_name)
  for per_job in (True, False):
    out, ref = _check_fused(mods, inp, conv=conv, mask_frames=1, per_job_bin=per_job)
    assert 0 < ref['depth'][0, 0, 0, 0] < 1   # the owner pixel holds a real rejected depth, not 0 and not the fill
