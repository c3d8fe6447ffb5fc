"""the bilinear resampling functions of utils/pano_utils.py
(rotate_pano, project_perspective_image, get_perspective_from_equirectangular_image,
equirectangular_pixel_rays, get_world_to_image_transform, crop_pano, tfa interpolate_bilinear).

CPU: properties of the numpy oracle.  GPU: the CUDA gather (se3ds_interpolate_bilinear) is
bit-exact against the oracle on identical query points; the full functions agree to 1e-4 absolute
(their query coordinates come from different sin / cos / atan2 implementations).
"""
import math

import numpy as np
import pytest

from oracle import ref_numpy as R

F32 = np.float32


def test_oracle_interpolate_bilinear_properties():
  rng = np.random.default_rng(0)
  grid = rng.uniform(0, 1, (2, 5, 7, 3)).astype(F32)
  ys, xs = np.meshgrid(np.arange(5), np.arange(7), indexing='ij')
  q = np.stack([ys, xs], -1).reshape(1, -1, 2).astype(F32).repeat(2, 0)
  # integer queries reproduce the grid (the last row / column goes through alpha = 1: a*(b-a)+a rounds)
  np.testing.assert_allclose(R.interpolate_bilinear(grid, q).reshape(2, 5, 7, 3), grid, rtol=0, atol=1e-7)
  np.testing.assert_allclose(R.interpolate_bilinear(grid, q[..., ::-1], indexing='xy').reshape(2, 5, 7, 3), grid, rtol=0, atol=1e-7)
  far = np.array([[[-3.0, -3.0], [99.0, 99.0], [2.5, 3.5]]], F32).repeat(2, 0)  # clamped outside
  out = R.interpolate_bilinear(grid, far)
  np.testing.assert_array_equal(out[:, 0], grid[:, 0, 0])
  np.testing.assert_array_equal(out[:, 1], grid[:, 4, 6])
  np.testing.assert_allclose(out[:, 2], grid[:, 2:4, 3:5].mean(axis=(1, 2)), rtol=1e-6)


def test_oracle_rotate_pano_identity_and_yaw():
  rng = np.random.default_rng(1)
  pano = rng.uniform(0, 1, (1, 16, 32, 3)).astype(F32)
  out = R.rotate_pano(pano, np.eye(3, dtype=F32)[None])
  # identity except the pole rows and the seam column (the reference's own ray table)
  np.testing.assert_allclose(out[:, 1:-1, 1:-1], pano[:, 1:-1, 1:-1], atol=2e-5)
  out2 = R.rotate_pano(pano, np.eye(3, dtype=F32)[None], output_height=8)
  assert out2.shape == (1, 8, 16, 3)
  with pytest.raises(ValueError):
    R.rotate_pano(pano[:, :, :30], np.eye(3, dtype=F32)[None])


def test_oracle_perspective_round_trip():
  """perspective -> equirect -> perspective reproduces the interior of a smooth image."""
  h, w = 48, 64
  yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing='ij')
  img = np.stack([xx, yy, 0.5 * (xx + yy)], -1).astype(F32)
  fov = [2 * math.atan(0.5 * (h - 1) / 40.0), 2 * math.atan(0.5 * (w - 1) / 40.0)]
  eq = R.project_perspective_image(img, fov, 256)
  k = np.array([[40, 0, 0.5 * (w - 1)], [0, 40, 0.5 * (h - 1)], [0, 0, 1]], F32)
  back = R.get_perspective_from_equirectangular_image(eq, k, np.eye(3, dtype=F32), h, w)
  assert back.shape == (h, w, 3)
  np.testing.assert_allclose(back[4:-4, 4:-4], img[4:-4, 4:-4], atol=0.02)


def test_oracle_world_to_image_and_crop():
  t = R.get_world_to_image_transform((48, 64), [1.0, 1.2])
  assert t.shape == (3, 3) and t[2, 2] == 1 and abs(t[0, 2] - 31.5) < 1e-6
  t2 = R.get_world_to_image_transform((48, 64), [1.0, 1.2], rotations=[0.0, 0.0])
  np.testing.assert_allclose(t, t2, atol=1e-6)
  pano = np.zeros((2, 64, 128, 3), np.int32)
  assert R.crop_pano(pano).shape == (2, 48, 128, 3)
  with pytest.raises(ValueError):
    R.crop_pano(np.zeros((4,)))


def test_oracle_antialiased_resize():
  rng = np.random.default_rng(5)
  img = rng.uniform(0, 255, size=(2, 24, 32, 3)).astype(np.float32)
  # same size: identity
  np.testing.assert_array_equal(R.tf_resize_antialias_triangle(img, (24, 32)), img)
  # enlargement: kernel scale 1, normalised triangle weights == bilinear weights
  up = R.tf_resize_antialias_triangle(img, (32, 32))
  np.testing.assert_allclose(up, R.tf_resize(img, (32, 32), 'bilinear'), rtol=0, atol=255 * 2e-6)
  # reduction: a real low-pass (constant stays constant, mean preserved, differs from plain bilinear)
  const = np.full((1, 32, 64, 1), 7.0, np.float32)
  np.testing.assert_allclose(R.tf_resize_antialias_triangle(const, (8, 16)), 7.0, rtol=1e-6)
  down = R.tf_resize_antialias_triangle(img, (12, 16))
  assert abs(down.mean() - img.mean()) < 0.5
  assert np.abs(down - R.tf_resize(img, (12, 16), 'bilinear')).max() > 1.0
  # crop_pano back to the original size, dtype restored by truncation
  pano = rng.integers(0, 256, size=(2, 64, 128, 3)).astype(np.uint8)
  back = R.crop_pano(pano, resize_to_original=True)
  assert back.shape == pano.shape and back.dtype == np.uint8
  near = R.crop_pano(pano, method='nearest', resize_to_original=True)
  assert set(np.unique(near)) <= set(np.unique(pano[:, 8:56]))


# ------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def gp():
  import torch
  from se3ds_b200.utils import pano_utils
  assert torch.cuda.is_available()
  return torch, pano_utils


@pytest.mark.gpu
@pytest.mark.parametrize('indexing', ['ij', 'xy'])
def test_cuda_interpolate_bilinear_bit_exact(gp, indexing):
  torch, pano = gp
  rng = np.random.default_rng(2)
  grid = rng.uniform(-1, 1, (3, 37, 53, 4)).astype(F32)
  q = rng.uniform(-4, 60, (3, 5000, 2)).astype(F32)
  q[:, :50] = np.round(q[:, :50])
  got = pano.interpolate_bilinear(torch.as_tensor(grid), torch.as_tensor(q), indexing=indexing).cpu().numpy()
  np.testing.assert_array_equal(got, R.interpolate_bilinear(grid, q, indexing=indexing))


@pytest.mark.gpu
def test_cuda_pixel_rays_kat(gp):
  """utils/pano_utils_test.py:35-65 through the CUDA-side shim."""
  torch, pano = gp
  rays = pano.equirectangular_pixel_rays(3).cpu().numpy()
  np.testing.assert_allclose(rays, R.equirectangular_pixel_rays(3), atol=1e-6)
  assert rays.shape == (3, 18)


@pytest.mark.gpu
def test_cuda_rotate_pano(gp):
  torch, pano = gp
  rng = np.random.default_rng(3)
  img = rng.uniform(0, 1, (2, 64, 128, 3)).astype(F32)
  a = rng.standard_normal((2, 3, 3))
  qm, _ = np.linalg.qr(a)
  qm = qm.astype(F32)
  got = pano.rotate_pano(torch.as_tensor(img), torch.as_tensor(qm)).cpu().numpy()
  want = R.rotate_pano(img, qm)
  assert got.shape == want.shape == (2, 64, 128, 3)
  # coordinates differ by ulps between libm and CUDA: tiny output differences, rare seam flips
  assert np.mean(np.abs(got - want) > 1e-4) < 2e-3
  small = pano.rotate_pano(torch.as_tensor(img), torch.as_tensor(qm), output_height=32)
  assert tuple(small.shape) == (2, 32, 64, 3)
  with pytest.raises(ValueError):
    pano.rotate_pano(torch.as_tensor(img[:, :, :100]), torch.as_tensor(qm))


@pytest.mark.gpu
@pytest.mark.parametrize('pad_mode,nearest', [('constant', False), ('mean', False), ('reflect', False), ('constant', True)])
def test_cuda_project_perspective_image(gp, pad_mode, nearest):
  torch, pano = gp
  rng = np.random.default_rng(4)
  img = rng.uniform(0, 1, (48, 64, 3)).astype(F32)
  fov = [1.0, 1.25]
  got = pano.project_perspective_image(torch.as_tensor(img), fov, 128, rotations=[0.1, -0.3], pad_mode=pad_mode,
                                       pad_value=0.25, round_to_nearest=nearest).cpu().numpy()
  want = R.project_perspective_image(img, fov, 128, rotations=[0.1, -0.3], pad_mode=pad_mode, pad_value=0.25,
                                     round_to_nearest=nearest)
  assert got.shape == want.shape == (128, 256, 3)
  assert np.mean(np.abs(got - want) > 1e-4) < (2e-2 if nearest else 2e-3)


@pytest.mark.gpu
def test_cuda_perspective_from_equirect(gp):
  torch, pano = gp
  rng = np.random.default_rng(5)
  eq = rng.uniform(0, 1, (128, 256, 3)).astype(F32)
  k = np.array([[60, 0, 39.5], [0, 60, 29.5], [0, 0, 1]], F32)
  a = rng.standard_normal((3, 3))
  qm, _ = np.linalg.qr(a)
  got = pano.get_perspective_from_equirectangular_image(torch.as_tensor(eq), k, qm.astype(F32), 60, 80).cpu().numpy()
  want = R.get_perspective_from_equirectangular_image(eq, k, qm.astype(F32), 60, 80)
  assert got.shape == (60, 80, 3)
  assert np.mean(np.abs(got - want) > 1e-4) < 2e-3


@pytest.mark.gpu
def test_cuda_crop_and_transform(gp):
  torch, pano = gp
  x = torch.zeros((2, 64, 128, 3), dtype=torch.int32)
  assert tuple(pano.crop_pano(x).shape) == (2, 48, 128, 3)
  rng = np.random.default_rng(9)
  for dtype, method in ((np.float32, 'bilinear'), (np.uint8, 'bilinear'), (np.uint8, 'nearest'), (np.int32, 'nearest')):
    src = rng.uniform(0, 255, size=(2, 64, 128, 3)).astype(dtype)
    got = pano.crop_pano(torch.from_numpy(src).cuda(), method=method, resize_to_original=True)
    want = R.crop_pano(src, method=method, resize_to_original=True)
    assert tuple(got.shape) == src.shape and got.cpu().numpy().dtype == dtype
    if method == 'nearest':
      np.testing.assert_array_equal(got.cpu().numpy(), want)
    elif dtype == np.float32:
      np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=255 * 2e-6)
    else:          # truncation can flip where the float result sits within rounding of an integer
      diff = np.abs(got.cpu().numpy().astype(np.int32) - want.astype(np.int32))
      assert diff.max() <= 1 and np.mean(diff > 0) < 1e-3
  single = pano.crop_pano(torch.from_numpy(src[0]).cuda(), method='nearest', resize_to_original=True)
  assert tuple(single.shape) == src.shape[1:]
  t = pano.get_world_to_image_transform((48, 64), [1.0, 1.2], rotations=[0.2, 0.4]).numpy()
  np.testing.assert_allclose(t, R.get_world_to_image_transform((48, 64), [1.0, 1.2], rotations=[0.2, 0.4]), atol=1e-5)
