"""Randomised parity sweep on the GPU: random shapes / conventions / knobs against the bit-exact oracle.
Usage: python tests/tools/gpu_fuzz.py [seconds] [seed]   (test tooling: it checks the CUDA path against oracle/)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from oracle import ref_exact as X, ref_numpy as R
from se3ds_b200 import _lib, guidance as g, synth


def oracle(inp, conv, mask_frames, per_job_bin, rot=None):
  rgb = inp['rgb'].astype(np.int32).copy()
  for k in range(mask_frames):
    rgb[:, k] = R.mask_pano(rgb[:, k], masked_region_value=-1)
  return X.reproject(rgb, inp['depth'], inp['src_pos'], inp['tgt_pos'], unproject_void=conv.unproject_void,
                     project_void=conv.project_void, mask_first_frame=False, per_job_bin=per_job_bin, tgt_rot=rot)


def rotations(rng, n, p):
  qm, _ = np.linalg.qr(rng.standard_normal((n * p, 3, 3)))
  qm *= np.sign(np.linalg.det(qm))[:, None, None]
  return qm.reshape(n, p, 3, 3).astype(np.float32)


def main(budget=40.0, seed=0):
  rng = np.random.default_rng(seed)
  t0, cases = time.time(), 0
  while time.time() - t0 < budget:
    h = int(rng.choice([3, 4, 5, 8, 16, 31, 32, 33, 64, 65, 96, 128, 130, 256]))
    n, s, p = int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.integers(1, 4))
    if h >= 128:
      n, s, p = min(n, 2), min(s, 2), min(p, 2)
    dist = str(rng.choice(['room', 'rand']))
    conv = [g.GAN_MANAGER, g.EVAL_METRIC][int(rng.integers(0, 2))]   # (SE3DS_MODEL's compaction has its own oracle in the tests)
    mask_frames = int(rng.integers(0, s + 1))
    per_job = bool(rng.integers(0, 2))
    lanes = int(rng.integers(1, 5))
    chunk_jobs = int(rng.integers(1, 4))
    inp = synth.make_inputs(n, s, p, h, seed=int(rng.integers(0, 1 << 30)), dist=dist, sweep=bool(rng.integers(0, 2)))
    if rng.integers(0, 4) == 0:  # degenerate depths: zeros, ones, out-of-range, a NaN
      d = inp['depth']
      d[rng.random(d.shape) < 0.2] = rng.choice([0.0, 1.0, -0.5, 1.5, np.nan])
    if rng.integers(0, 4) == 0:  # int32 colours with void values in them (the generic, non-FAST kernels)
      rgb = inp['rgb'].astype(np.int32)
      rgb[rng.random(rgb.shape) < 0.05] = -1
      rgb[rng.random(rgb.shape[:-1]) < 0.03] = conv.unproject_void
      inp['rgb'] = rgb
    rot = rotations(rng, n, p) if rng.integers(0, 4) == 0 else None
    ws = _lib.Workspace(0, 0, h * 2 * h * (16 + 8 * s) * chunk_jobs * lanes)
    ws.lanes(lanes, 1, 1)
    t = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
    trot = None if rot is None else torch.as_tensor(rot).cuda()
    want = oracle(inp, conv, mask_frames, per_job, rot)
    for key64, winner in ((False, False), (True, False), (True, True)):
      out = g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=mask_frames,
                        unproject_void=conv.unproject_void, project_void=conv.project_void, filter_void=False,
                        per_job_bin=per_job, return_winner=winner, workspace=ws, key64=key64, tgt_rot=trot)
      torch.cuda.synchronize()
      tag = (h, n, s, p, dist, conv.unproject_void, conv.project_void, mask_frames, per_job, lanes, chunk_jobs, key64, winner,
             str(inp['rgb'].dtype), rot is not None)
      for name, ref in (('proj_depth', 'depth'), ('proj_mask', 'mask'), ('proj_image', 'image')):
        a, b = out[name].cpu().numpy(), want[ref]
        if not np.array_equal(a, b, equal_nan=True):
          bad = np.argwhere(a != b)[:8]
          os.makedirs('gpurun_out', exist_ok=True)
          np.savez('gpurun_out/fuzz_fail.npz', **inp, rot=(rot if rot is not None else np.zeros(0)), params=np.array([h, n, s, p, conv.unproject_void, conv.project_void, mask_frames, int(per_job), lanes, chunk_jobs, int(key64), int(winner)]))
          raise AssertionError((name, tag, int(np.sum(a != b)), bad.tolist(), [(float(a[tuple(i)]), float(b[tuple(i)])) for i in bad]))
      if winner:
        assert np.array_equal(out['winner'].cpu().numpy(), want['winner']), ('winner', tag)
    # compact outputs + expand == the float32 contract; a frame ring with spare capacity == the dense call
    outc = g.expand_guidance(g.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=mask_frames,
                                         unproject_void=conv.unproject_void, project_void=conv.project_void, per_job_bin=per_job,
                                         workspace=ws, tgt_rot=trot, compact=True))
    pad = int(rng.integers(0, 3))
    ring = {k: torch.cat([t[k], t[k][:, :1].expand(-1, pad, *t[k].shape[2:])], dim=1).contiguous() if pad else t[k]
            for k in ('rgb', 'depth', 'src_pos')}
    outr = g.reproject(ring['rgb'], ring['depth'], ring['src_pos'], t['tgt_pos'], mask_frames=mask_frames, frames=s,
                       unproject_void=conv.unproject_void, project_void=conv.project_void, per_job_bin=per_job,
                       workspace=ws, tgt_rot=trot)
    for name, ref in (('proj_depth', 'depth'), ('proj_mask', 'mask'), ('proj_image', 'image')):
      assert np.array_equal(outc[name].cpu().numpy(), want[ref], equal_nan=True), ('compact', name, tag)
      assert np.array_equal(outr[name].cpu().numpy(), want[ref], equal_nan=True), ('ring', name, tag)
    ws.close()
    cases += 1
  print(f'FUZZ OK: {cases} random cases x 3 key modes bit-identical to the oracle in {time.time() - t0:.0f} s (seed {seed})')


if __name__ == '__main__':
  main(float(sys.argv[1]) if len(sys.argv) > 1 else 40.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
