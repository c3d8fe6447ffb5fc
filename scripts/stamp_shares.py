"""Per-kernel shares of the pipelined c2 step from the end-of-kernel %globaltimer stamps (profile mode 2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from se3ds_b200 import _lib, guidance, synth

dist = sys.argv[1] if len(sys.argv) > 1 else 'room'
torch.cuda.set_device(0)
ws = _lib.Workspace(0)
plans = []
for r in range(4):
  inp = synth.make_inputs(8, 1, 1, 512, seed=r, dist=dist)
  t = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
  plans.append(guidance.prepare(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], mask_frames=1, workspace=ws, inputs_ready=True))
for i in range(50):
  plans[i % 4].run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(1000):
  plans[i % 4].run()
e1.record(); torch.cuda.synchronize()
print('step us %.2f' % (e0.elapsed_time(e1)))
ws.profile(2)
for i in range(1001):
  plans[i % 4].run()
ms, n = ws.profile_read_stamps()
print('stamps: chunks', n, 'K2 %.2f K3 %.2f K4 %.2f us; sum %.2f' % tuple([m / n * 1e3 for m in ms] + [sum(ms) / n * 1e3]))
ws.profile(0)
