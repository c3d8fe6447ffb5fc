"""Prints the certified fraction of the fast projection for the bench workloads (verify mode)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se3ds_b200 import _lib, guidance, synth

for dist in ('room', 'rand'):
  for h in (512,):
    inp = synth.make_inputs(8, 1, 1, h, seed=1000, dist=dist)
    t = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
    for scale in (1e-6, 2.5e-7):
      ws = _lib.Workspace(0)
      ws.projection_mode(2, scale)
      guidance.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, workspace=ws)
      v = ws.verify_read()
      print(dist, h, 'margin_scale', scale, 'certified %.4f' % (v['certified'] / v['points']), v, flush=True)
      ws.close()
