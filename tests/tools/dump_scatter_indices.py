"""Writes scripts/micro/bin/idx_{room,rand}.i32: the target pixel of every source point of one bench pano as
the oracle computes it (-1 = rejected), the "real" address patterns of scripts/micro/scatter_micro.cu."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import ref_exact as E, ref_numpy as R
from se3ds_b200 import synth
H, W = 512, 1024
for dist in ('room', 'rand'):
  inp = synth.make_inputs(1, 1, 1, H, seed=0, dist=dist)
  rgb = inp['rgb'][0, 0].astype(np.int32)
  rgbm = R.mask_pano(rgb[None], 0.125, -1)
  xyz1, feats = E.equirectangular_to_pointcloud(rgbm, inp['depth'][0:1, 0], -1, 20.0)
  xyz1 = xyz1.copy()
  xyz1[:, :3] = (xyz1[:, :3] + inp['src_pos'][0, 0][None, :, None]) - inp['tgt_pos'][0, 0][None, :, None]
  out = E.splat(xyz1, feats, H, W, 20.0, -1.0, 0.0, 0)
  flat = out['flat'][0].astype(np.int32).copy()
  valid = out['valid'][0].astype(bool)
  flat[~valid] = -1
  os.makedirs(os.path.join(ROOT, 'scripts', 'micro', 'bin'), exist_ok=True)
  flat.tofile(os.path.join(ROOT, 'scripts', 'micro', 'bin', f'idx_{dist}.i32'))
  print(dist, valid.mean())
