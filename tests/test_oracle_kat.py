"""Pins the numpy oracle against the reference's own known-answer tests.

Each test restates one test of the reference (file:line in the docstring) with
the oracle in place of the TF functions.  CPU only.
"""
import math

import numpy as np
import pytest

from oracle import ref_numpy as R

F32 = np.float32


def test_tf_linspace_endpoints_and_interior():
  v = R.tf_linspace(0.0, 1.0, 5)
  assert v.dtype == np.float32
  np.testing.assert_array_equal(v, np.array([0, 0.25, 0.5, 0.75, 1.0], F32))
  e, h = R.equirect_angle_tables(4, 8)
  assert e[0] == F32(0.5 * np.pi / 4) and e[-1] == F32(np.pi - 0.5 * np.pi / 4)
  assert h[0] == F32(1.5 * np.pi - 0.5 * np.pi / 4)


def test_equirectangular_pixel_rays_kat():
  """utils/pano_utils_test.py:35-65 (3x6x3 ray table, assertAllClose 1e-6)."""
  rays = R.equirectangular_pixel_rays(3)
  rays = rays.T.reshape(3, 6, 3)
  expected = np.array([
      [[0.0, -1.0, 0.0]] * 6,
      [[0.0, 0.0, -1.0],
       [-9.5105648e-01, 4.3711388e-08, -3.0901703e-01],
       [-5.8778524e-01, 4.3711388e-08, 8.0901694e-01],
       [5.8778524e-01, 4.3711388e-08, 8.0901694e-01],
       [9.5105648e-01, 4.3711388e-08, -3.0901703e-01],
       [0.0, 0.0, -1.0]],
      [[0.0, 1.0, 0.0]] * 6], F32)
  np.testing.assert_allclose(rays, expected, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('batch_size,image_size', [(2, 64), (1, 128)])
def test_feats_to_equirectangular(batch_size, image_size):
  """utils/pano_utils_test.py:67-87."""
  rng = np.random.default_rng(0)
  m = image_size**2
  feats = rng.integers(0, R.NUM_MP3D_CLASSES, (batch_size, m)).astype(np.int32)
  xyz = rng.standard_normal((batch_size, 3, m)).astype(F32)
  xyz1 = np.concatenate([xyz, np.ones((batch_size, 1, m), F32)], axis=1)
  d, f = R.project_feats_to_equirectangular(feats, xyz1, image_size, image_size * 2,
                                            R.INVALID_SEM_VALUE, R.DEPTH_SCALE)
  assert d.shape == (batch_size, image_size, image_size * 2)
  assert f.shape == (batch_size, image_size, image_size * 2)
  assert d.min() >= 0 and d.max() <= 1
  assert f.min() >= 0 and f.max() <= R.NUM_MP3D_CLASSES


@pytest.mark.parametrize('batch_size,image_size,multi', [(2, 64, False), (1, 128, False),
                                                         (2, 64, True), (1, 128, True)])
def test_filter_equirectangular(batch_size, image_size, multi):
  """utils/pano_utils_test.py:89-111."""
  rng = np.random.default_rng(1)
  shape = (batch_size, image_size, 2 * image_size) + ((3,) if multi else ())
  feats = rng.integers(0, R.NUM_MP3D_CLASSES, shape).astype(np.int32)
  depth = rng.uniform(0, R.DEPTH_SCALE, (batch_size, image_size, 2 * image_size)).astype(F32)
  xyz1, ff = R.equirectangular_to_pointcloud(feats, depth, R.INVALID_SEM_VALUE, R.DEPTH_SCALE)
  assert xyz1.shape == (batch_size, 4, 2 * image_size**2)
  assert ff.shape == (batch_size, 2 * image_size**2) + ((3,) if multi else ())
  assert ff.min() >= 0 and ff.max() <= R.NUM_MP3D_CLASSES


@pytest.mark.parametrize('batch_size,h,dtype', [(2, 64, np.float32), (2, 64, np.int32), (1, 256, np.int32)])
def test_mask_pano(batch_size, h, dtype):
  """utils/pano_utils_test.py:113-123."""
  rng = np.random.default_rng(2)
  pano = rng.uniform(0, 255, (batch_size, h, 2 * h, 3)).astype(dtype)
  out = R.mask_pano(pano)
  assert out.shape == pano.shape and out.dtype == pano.dtype
  assert np.all(out[:, 0] == 0) and np.all(out[:, -1] == 0)
  mh = int(h * 0.125)
  # the <= at pano_utils.py:263 keeps row H - mh
  np.testing.assert_array_equal(out[:, mh:h - mh + 1], pano[:, mh:h - mh + 1])
  assert np.all(out[:, h - mh + 1:] == 0)


def test_equirect_void_dtype_errors():
  """utils/pano_utils.py:190-197."""
  with pytest.raises(ValueError):
    R.equirectangular_to_pointcloud(np.zeros((1, 4, 8, 3, 1), np.int32), np.zeros((1, 4, 8), F32), 0, 20.0)
  with pytest.raises(ValueError):
    R.equirectangular_to_pointcloud(np.zeros((1, 4, 8, 3), np.uint8), np.zeros((1, 4, 8), F32), -1, 20.0)
  with pytest.raises(AssertionError):
    R.equirectangular_to_pointcloud(np.zeros((1, 4, 9, 3), np.int32), np.zeros((1, 4, 9), F32), -1, 20.0)


@pytest.mark.parametrize('batch_size,image_size', [(2, 64), (1, 128)])
def test_filtered_coords_and_feats(batch_size, image_size):
  """utils/point_cloud_utils_test.py:26-40."""
  rng = np.random.default_rng(3)
  feats = rng.integers(0, R.NUM_MP3D_CLASSES, (batch_size, image_size, image_size)).astype(np.int32)
  depth = rng.uniform(0, R.DEPTH_SCALE, (batch_size, image_size, image_size)).astype(F32)
  xyz1, ff = R.get_filtered_coords_and_feats(feats, depth, R.DEPTH_SCALE)
  assert xyz1.shape == (batch_size, 4, image_size * image_size)
  assert ff.shape == (batch_size, image_size * image_size)
  assert ff.min() >= 0 and ff.max() <= R.NUM_MP3D_CLASSES


@pytest.mark.parametrize('batch_size,image_size,multi', [(2, 64, False), (1, 128, False),
                                                         (2, 64, True), (1, 128, True)])
def test_project_to_feat(batch_size, image_size, multi):
  """utils/point_cloud_utils_test.py:42-64 (square, non-equirect target)."""
  rng = np.random.default_rng(4)
  shape = (batch_size, image_size, image_size) + ((3,) if multi else ())
  feats = rng.integers(0, R.NUM_MP3D_CLASSES, shape).astype(np.int32)
  depth = rng.uniform(0, R.DEPTH_SCALE, (batch_size, image_size, image_size)).astype(F32)
  xyz1, ff = R.get_filtered_coords_and_feats(feats, depth, R.DEPTH_SCALE)
  pd, pf = R.project_to_feat(xyz1, ff, image_size, image_size, R.DEPTH_SCALE, R.INVALID_SEM_VALUE)
  assert pd.shape == (batch_size, image_size, image_size)
  assert pd.min() >= 0 and pd.max() <= 1
  assert pf.shape == shape
  assert pf.min() >= feats.min() and pf.max() <= feats.max()


@pytest.mark.parametrize('batch_size,image_size', [(1, 128), (2, 128), (1, 256)])
def test_identity_reprojection(batch_size, image_size):
  """models/models_test.py:38-68: add a pano at p, project at p => >= 95 % RGB equal."""
  rng = np.random.default_rng(5)
  rgb = rng.integers(0, 255, (batch_size, image_size, image_size * 2, 3)).astype(np.uint8)
  sem = rng.integers(0, R.NUM_MP3D_CLASSES, (batch_size, image_size, image_size * 2, 1)).astype(np.uint8)
  depth = rng.uniform(0, 1, (batch_size, image_size, image_size * 2)).astype(F32)
  pos = rng.standard_normal((batch_size, 3)).astype(F32)
  model = R.SE3DSMemoryOracle(image_size, batch_size=batch_size)
  model.add_to_memory(rgb, sem, depth, pos, mask_blurred=False)
  out = model.project(pos)
  proj_rgb_u8 = (out['proj_image'] * 255).astype(np.uint8)
  rgb_equal = np.all(proj_rgb_u8 == rgb, axis=-1)
  assert rgb_equal.mean() >= 0.95
  assert out['proj_image'].shape == (batch_size, image_size, image_size * 2, 3)
  assert out['proj_mask'].shape == (batch_size, image_size, image_size * 2, 1)
  assert out['proj_depth'].min() >= 0 and out['proj_depth'].max() <= 1


def test_internal_point_cloud_representation():
  """models/models_test.py:81-137: plane KAT (24 columns; y == 1 / x == 2)."""
  batch_size, image_size = 2, 4
  rng = np.random.default_rng(6)
  rgb = rng.integers(0, 255, (batch_size, image_size, image_size * 2, 3)).astype(np.uint8)
  sem = rng.integers(1, R.NUM_MP3D_CLASSES, (batch_size, image_size, image_size * 2, 1)).astype(np.uint8)
  offset = 0.5 * np.pi / image_size
  heading = R.tf_linspace(-np.pi + offset, np.pi - offset, image_size * 2)
  pitch = R.tf_linspace(0.5 * np.pi - offset, -0.5 * np.pi + offset, image_size)
  x_depth = (F32(1.0) / np.cos(heading))[None, :]
  depth = x_depth / np.cos(pitch)[:, None]
  depth = np.where(depth > 0, depth, 0).astype(F32)
  depth1 = np.roll(depth, image_size // 2, -1)
  test_depth = np.stack([depth, depth1], axis=0) / F32(R.DEPTH_SCALE)
  model = R.SE3DSMemoryOracle(image_size, batch_size=batch_size)
  start = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]], F32)
  model.add_to_memory(rgb, sem, test_depth, start, mask_blurred=False)
  mem = model.get_memory_state()
  pc = mem.rgb_coords
  assert pc.shape == (batch_size, 4, 24)
  for ix, (axis, value) in enumerate([(1, 1), (0, 2)]):
    valid = np.any(mem.rgb[ix] != R.INVALID_RGB_VALUE, axis=1)
    filtered = pc[ix][:, valid]
    assert filtered.shape[1] == image_size**2
    np.testing.assert_allclose(filtered[axis], np.full(image_size**2, value, F32), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('distance,depth_distance,expected', [(0.5, 0.5, 1.0), (0.3, 0.5, 0.0)])
def test_proportion_invalid(distance, depth_distance, expected):
  """inference/perturbation_utils_test.py:30-41."""
  h, w = 64, 128
  depth = np.full((h, w), depth_distance / R.DEPTH_SCALE, F32)
  assert R.get_proportion_invalid_for_depth(np.array([0.0, distance, 0.0], F32), depth) == expected


@pytest.mark.parametrize('offset,centre', [([0.0, 0.5, 0.0], (0.5, 0.5)), ([0.5, 0.5, 0.0], (0.75, 0.75))])
def test_proportion_invalid_offsets(offset, centre):
  """inference/perturbation_utils_test.py:43-94 (forward and diagonal)."""
  h, w, pad = 64, 128, 10
  img = np.full((h, w), 1.0, F32)
  hs, ws = int(h * centre[0]), int(w * centre[1])
  img[hs - pad:hs + pad, ws - pad:ws + pad] = 0.0
  assert R.get_proportion_invalid_for_depth(np.array(offset, F32), img) > 0.0
  img = np.full((h, w), 1.0, F32)
  img[:pad, :pad] = 0.0
  assert R.get_proportion_invalid_for_depth(np.array(offset, F32), img) == 0.0


def test_reject_bin_lands_on_pixel_zero():
  """utils/point_cloud_utils.py:150-153,168-176: rejected points all scatter to flat index 0."""
  rng = np.random.default_rng(7)
  h = 32
  rgb = rng.integers(0, 255, (2, 1, h, 2 * h, 3)).astype(np.uint8)
  depth = rng.uniform(0.02, 0.6, (2, 1, h, 2 * h)).astype(F32)
  src = np.zeros((2, 1, 3), F32)
  tgt = np.array([[1.0, 0.3, 0.05], [0.7, -0.2, 0.0]], F32)
  image, d, mask, dbg = R.reproject_trajectory(rgb, depth, src, tgt, mask_first_frame=False)
  rejected = dbg['flat_kept'].reshape(-1) == 0
  feats = rgb.reshape(-1, 3).astype(F32)
  expect = np.maximum(feats[rejected].max(axis=0), 0)
  np.testing.assert_array_equal(dbg['raw_rgb'][0, 0, 0], expect)
  # batch item 1's pixel (0,0) only sees genuine hits
  assert dbg['raw_rgb'][0, 0, 0].max() >= 250


def test_tf_resize_semantics():
  """tf.image.resize with half-pixel centres: x2 nearest replicates, identity bilinear is exact,
  x2 bilinear interpolates between neighbours and clamps at the border."""
  rng = np.random.default_rng(9)
  img = rng.integers(0, 255, (1, 4, 6, 2)).astype(np.int32)
  up = R.tf_resize(img, (8, 12), 'nearest')
  assert up.dtype == np.int32
  np.testing.assert_array_equal(up[:, ::2, ::2], img)
  np.testing.assert_array_equal(up[:, 1::2, 1::2], img)
  np.testing.assert_array_equal(R.tf_resize(img, (4, 6), 'bilinear'), img.astype(F32))
  row = np.array([0.0, 4.0, 8.0], F32).reshape(1, 1, 3, 1)
  out = R.tf_resize(row, (1, 6), 'bilinear')[0, 0, :, 0]
  np.testing.assert_allclose(out, [0, 1, 3, 5, 7, 8], atol=1e-6)
  down = R.tf_resize(img, (2, 3), 'nearest')
  np.testing.assert_array_equal(down, img[:, 1::2, 1::2])
  xyz1, ff = R.equirectangular_to_pointcloud(img[:, :3, :, 0], np.full((1, 3, 6), 0.5, F32), 0, 20.0, size_mult=2.0)
  assert xyz1.shape == (1, 4, 72) and ff.shape == (1, 72)
