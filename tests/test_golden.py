"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them (no silent drift).  GPU: the CUDA path reproduces them."""
import glob
import os

import numpy as np
import pytest

from oracle import ref_exact as X

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', '*.npz')))


def test_golden_files_exist():
  assert len(GOLDEN) >= 4


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
  g = np.load(path)
  o = X.reproject(g['rgb'], g['depth'], g['src_pos'], g['tgt_pos'], unproject_void=int(g['unproject_void']),
                  project_void=int(g['project_void']), mask_first_frame=bool(g['mask_first_frame']),
                  per_job_bin=bool(g['per_job_bin']))
  for k, r in (('image', 'image'), ('proj_depth', 'depth'), ('mask', 'mask'), ('winner', 'winner'), ('flat', 'flat'), ('rad', 'rad')):
    np.testing.assert_array_equal(o[r], g[k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_reproduces_golden(path):
  import torch
  from se3ds_b200 import guidance
  g = np.load(path)
  t = {k: torch.as_tensor(g[k]).cuda() for k in ('rgb', 'depth', 'src_pos', 'tgt_pos')}
  out = guidance.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=int(bool(g['mask_first_frame'])),
                           unproject_void=int(g['unproject_void']), project_void=int(g['project_void']),
                           per_job_bin=bool(g['per_job_bin']), return_winner=True)
  np.testing.assert_array_equal(out['winner'].cpu().numpy(), g['winner'])
  np.testing.assert_array_equal(out['proj_depth'].cpu().numpy(), g['proj_depth'])
  np.testing.assert_array_equal(out['proj_mask'].cpu().numpy(), g['mask'])
  np.testing.assert_array_equal(out['proj_image'].cpu().numpy(), g['image'])
