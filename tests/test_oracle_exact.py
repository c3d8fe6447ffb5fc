"""Cross-checks the canonical-arithmetic oracle (oracle/ref_exact.c, the bit
authority the CUDA kernels are compared with) against the literal numpy
restatement (oracle/ref_numpy.py, the semantic authority) and against the
reference's known-answer tests.  CPU only.
"""
import numpy as np
import pytest

from oracle import ref_exact as X
from oracle import ref_numpy as R

F32 = np.float32


def _ulp_err(a32, truth64):
  ulp = np.spacing(np.abs(truth64.astype(F32))).astype(np.float64)
  return np.abs(a32.astype(np.float64) - truth64) / ulp


def test_canonical_atan2_accuracy():
  rng = np.random.default_rng(0)
  x = rng.standard_normal(2_000_000).astype(F32) * F32(5)
  y = rng.standard_normal(2_000_000).astype(F32) * F32(5)
  got = X.atan2f(y, x)
  err = _ulp_err(got, np.arctan2(y.astype(np.float64), x.astype(np.float64)))
  assert err.max() < 2.0, err.max()
  assert err.mean() < 0.5
  # axes and signs
  ax = np.array([0, 0, 1, -1, 0, 1, -1, 1, -1], F32)
  ay = np.array([0, 1, 0, 0, -1, 1, 1, -1, -1], F32)
  np.testing.assert_allclose(X.atan2f(ay, ax), np.arctan2(ay, ax), rtol=3e-7, atol=0)


def test_canonical_acos_accuracy():
  rng = np.random.default_rng(1)
  q = np.concatenate([rng.uniform(-1, 1, 2_000_000), [-1.0, -0.5, 0.0, 0.5, 1.0],
                      1 - np.logspace(-7, -1, 1000)]).astype(F32)
  got = X.acosf(q)
  err = _ulp_err(got, np.arccos(q.astype(np.float64)))
  assert err.max() < 2.5, err.max()
  assert err.mean() < 0.5
  assert X.acosf(np.array([1.0], F32))[0] == 0.0
  assert X.acosf(np.array([-1.0], F32))[0] == F32(np.pi)


def test_tables_match_numpy_within_one_ulp():
  for h in (4, 64, 512):
    elev, head, se, ce, sh, ch = X.tables(h, 2 * h)
    e_np, h_np = R.equirect_angle_tables(h, 2 * h)
    np.testing.assert_array_equal(elev, e_np)
    np.testing.assert_array_equal(head, h_np)
    for mine, theirs in ((se, np.sin(e_np)), (ce, np.cos(e_np)), (sh, np.sin(h_np)), (ch, np.cos(h_np))):
      # correctly rounded double -> float vs numpy's float32 kernels
      assert np.max(np.abs(mine.astype(np.float64) - theirs.astype(np.float64))) <= 1.2e-7


def test_unproject_matches_numpy():
  rng = np.random.default_rng(2)
  depth = rng.uniform(-0.1, 1.1, (2, 64, 128)).astype(F32)
  feats = rng.integers(0, 255, (2, 64, 128, 3)).astype(np.int32)
  xyz_x, f_x = X.equirectangular_to_pointcloud(feats, depth, -1, 20.0)
  xyz_n, f_n = R.equirectangular_to_pointcloud(feats, depth, -1, 20.0)
  np.testing.assert_array_equal(f_x, f_n)
  np.testing.assert_allclose(xyz_x, xyz_n, rtol=0, atol=5e-6)
  assert xyz_x.shape == (2, 4, 64 * 128)


@pytest.mark.parametrize('h,seed', [(64, 0), (128, 1), (256, 2)])
def test_splat_matches_numpy_up_to_boundary_cases(h, seed):
  """Pixel indices of the canonical pipeline differ from the libm pipeline only
  for 1-ulp boundary cases (<= 2e-4 of the points); where the indices agree for
  every point of a pixel, depth is within 1e-5 relative and features equal."""
  rng = np.random.default_rng(seed)
  n, w = 2, 2 * h
  rgb = rng.integers(0, 256, (n, 1, h, w, 3)).astype(np.int32)
  depth = rng.uniform(0, 1, (n, 1, h, w)).astype(F32)
  src = np.zeros((n, 1, 3), F32)
  tgt = np.array([[1.0, 0.3, 0.05], [-0.4, 0.8, -0.1]], F32)
  ox = X.reproject(rgb, depth, src, tgt, mask_first_frame=True)
  image, d, mask, dbg = R.reproject_trajectory(rgb, depth, src, tgt, mask_first_frame=True)
  flat_x, flat_n = ox['flat'], dbg['flat']
  disagree = np.mean(flat_x != flat_n)
  assert disagree <= 2e-4, disagree
  touched = np.zeros(n * h * w, bool)
  bad = flat_x != flat_n
  touched[flat_x[bad]] = True
  touched[flat_n[bad]] = True
  touched[0] = True  # the global reject bin collects every disagreement
  ok = ~touched.reshape(n, h, w)
  np.testing.assert_allclose(ox['depth'][..., 0][ok], d[..., 0][ok], rtol=1e-5, atol=1e-7)
  # tolerance test (d < dmin + 0.1) can flip for |delta| ~ 1 ulp: allow a tiny fraction
  feat_bad = np.any(ox['raw_rgb'][ok] != dbg['raw_rgb'][ok], axis=-1).mean()
  assert feat_bad <= 2e-4, feat_bad
  mask_bad = np.mean(ox['mask'][..., 0][ok] != mask[..., 0][ok])
  assert mask_bad <= 2e-4


def test_exact_identity_reprojection_and_plane_kat():
  """models/models_test.py:64-68 and :81-137 through the canonical arithmetic."""
  rng = np.random.default_rng(3)
  h = 128
  rgb = rng.integers(0, 255, (1, 1, h, 2 * h, 3)).astype(np.int32)
  depth = rng.uniform(0, 1, (1, 1, h, 2 * h)).astype(F32)
  pos = rng.standard_normal((1, 1, 3)).astype(F32)
  o = X.reproject(rgb, depth, pos, pos[:, 0], mask_first_frame=False)
  proj_u8 = (o['image'] * 255).astype(np.uint8)
  assert np.all(proj_u8[0] == rgb[0, 0], axis=-1).mean() >= 0.95
  # plane KAT
  s = 4
  offset = 0.5 * np.pi / s
  heading = R.tf_linspace(-np.pi + offset, np.pi - offset, 2 * s)
  pitch = R.tf_linspace(0.5 * np.pi - offset, -0.5 * np.pi + offset, s)
  d0 = (F32(1.0) / np.cos(heading))[None, :] / np.cos(pitch)[:, None]
  d0 = np.where(d0 > 0, d0, 0).astype(F32)
  test_depth = np.stack([d0, np.roll(d0, s // 2, -1)], 0) / F32(20.0)
  xyz1, valid = X.unproject(test_depth, 20.0)
  xyz1 = xyz1 + np.array([[0, 0, 0, 0], [1, 0, 0, 0]], F32)[:, :, None]
  assert valid.sum(axis=1).tolist() == [16, 16]
  np.testing.assert_allclose(xyz1[0, 1][valid[0]], 1.0, rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(xyz1[1, 0][valid[1]], 2.0, rtol=1e-6, atol=1e-6)
  assert np.any(valid, axis=0).sum() == 24


def test_winner_definition():
  """winner = lowest index among the valid points at the pixel minimum."""
  coords = np.zeros((1, 4, 4), F32)
  # four points straight ahead (+y), two of them at the same minimum depth
  coords[0, 1] = [2.0, 1.0, 1.0, 3.0]
  coords[0, 3] = 1
  feats = np.array([[[10, 0, 0], [20, 5, 0], [5, 30, 0], [99, 99, 99]]], F32)
  o = X.splat(coords, feats, 8, 16, 20.0, -1.0)
  pix = o['flat'][0, 0]
  assert np.all(o['flat'][0] == pix) and pix != 0
  assert o['winner'].reshape(-1)[pix] == 1
  assert o['zbuf'].reshape(-1)[pix] == 1.0
  # near-min set {1, 2} (depth 1.0 < 1.0 + 0.1); per-channel max -> (20, 30, 0)
  np.testing.assert_array_equal(o['feat'].reshape(-1, 3)[pix], [20, 30, 0])
  # points 0 and 3 are rejected -> bin at flat index 0 gets max(10,99)=99
  np.testing.assert_array_equal(o['feat'].reshape(-1, 3)[0], [99, 99, 99])
  assert o['kept'][0].tolist() == [0, pix, pix, 0]
