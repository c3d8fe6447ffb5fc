"""Host-link probe: raw pinned copies (one direction, both at once) and se3ds_reproject_host per batch size."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from se3ds_b200 import guidance, _lib

dev = torch.device('cuda', 0)
NB = 29360128
hb_in = torch.empty(NB, dtype=torch.uint8, pin_memory=True)
hb_out = torch.empty(NB, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(NB, dtype=torch.uint8, device=dev)
d_out = torch.empty(NB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
def timeit(fn, reps=20):
  fn(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(reps): fn()
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / reps * 1e3
def h2d():
  with torch.cuda.stream(s1): d_in.copy_(hb_in, non_blocking=True)
def d2h():
  with torch.cuda.stream(s2): hb_out.copy_(d_out, non_blocking=True)
def both():
  h2d(); d2h()
def chunks(k):
  def f():
    c = NB // k
    for i in range(k):
      with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(hb_in[i*c:(i+1)*c], non_blocking=True)
      with torch.cuda.stream(s2): hb_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
  return f
res['h2d_ms'] = timeit(h2d); res['d2h_ms'] = timeit(d2h); res['both_ms'] = timeit(both)
res['both_16chunks_ms'] = timeit(chunks(16))
rng = np.random.default_rng(0)
for n in (1, 2, 4, 8, 16):
  h, w = 512, 1024
  rgb = torch.from_numpy(rng.integers(0, 256, (n, 1, h, w, 3), dtype=np.uint8)).pin_memory()
  depth = torch.from_numpy(rng.uniform(0.05, 0.9, (n, 1, h, w)).astype(np.float32)).pin_memory()
  sp = torch.zeros(n, 1, 3).pin_memory(); tp = torch.full((n, 1, 3), 0.3).pin_memory()
  for compact in (True, False):
    out = {}
    ws = _lib.Workspace(0)
    f = lambda: guidance.reproject_host(rgb, depth, sp, tp, mask_frames=1, out=out, device=0, workspace=ws, compact=compact)
    res['host_n%d_%s_ms' % (n, 'compact' if compact else 'f32')] = timeit(f, 10)
print(json.dumps(res, indent=1))
