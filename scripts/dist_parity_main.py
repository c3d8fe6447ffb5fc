"""Run under torchrun on N GPUs: the sharded driver (NCCL all-gather + bin all-reduce) must equal
the single-call result of rank 0's own GPU, bit for bit, for every bin mode and job shape."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from se3ds_b200 import guidance, parallel, synth  # noqa: E402


def main():
  rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  ok = True
  multicast_ok, tested_multicast = True, False
  for (n, s, p, sweep) in ((8, 1, 1, False), (1, 1, 16, True), (3, 2, 5, True), (1, 1, 1, False)):
    inp = synth.make_inputs(n, s, p, 64, seed=5, dist='rand', sweep=sweep)
    t = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
    for mode in ('call', 'job'):
      ref = guidance.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], per_job_bin=(mode == 'job'),
                               return_winner=True)
      ref = {k: v.clone() for k, v in ref.items()}
      for wire, chunks in (('compact', 0), ('compact', 1), ('f32', 0)):   # pipelined compact gather, one piece, float32 wire
        got = parallel.reproject_sharded(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], bin_mode=mode,
                                         return_winner=True, wire=wire, chunks=chunks)
        for k in ('proj_image', 'proj_depth', 'proj_mask', 'winner'):
          same = torch.equal(got[k], ref[k])
          ok &= same
          if not same:
            print(f'rank {rank}: MISMATCH {k} n={n} s={s} p={p} mode={mode} wire={wire} chunks={chunks}', flush=True)
    if (n * p) % dist.get_world_size() == 0:   # the prepared (serving) form
      for mode in ('call', 'job'):
        ref = guidance.reproject(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], per_job_bin=(mode == 'job'))
        ref = {k: v.clone() for k, v in ref.items()}
        for wire, pieces in (('nccl', 1), ('nccl', 2), ('multicast', 1)):
          if wire == 'multicast' and not multicast_ok:
            continue
          try:
            plan = parallel.ShardedReprojection(t['rgb'], t['depth'], t['src_pos'], t['tgt_pos'], bin_mode=mode, pieces=pieces, wire=wire)
          except Exception as e:  # pylint: disable=broad-except
            if wire != 'multicast':
              raise
            multicast_ok = False   # no NVSwitch multicast on this box: said once, not a parity failure
            print(f'rank {rank}: multicast wire unavailable: {type(e).__name__}: {e}', flush=True)
            continue
          for _ in range(3):
            got = plan.run()
          for k in ('proj_image', 'proj_depth', 'proj_mask'):
            same = torch.equal(got[k], ref[k])
            ok &= same
            if not same:
              print(f'rank {rank}: MISMATCH prepared {k} n={n} s={s} p={p} mode={mode} wire={wire} pieces={pieces}', flush=True)
          if wire == 'multicast':
            tested_multicast = True
  flag = torch.tensor([int(ok)], device='cuda')
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  if rank == 0:
    print('DIST PARITY', 'OK' if flag.item() == 1 else 'FAILED', 'world', dist.get_world_size(),
          'multicast wire', 'tested' if tested_multicast else 'NOT available', flush=True)
  dist.destroy_process_group()
  sys.exit(0 if flag.item() == 1 else 1)


if __name__ == '__main__':
  main()
