"""torchrun -n N scripts/sharded_probe.py: the strong-scaling sharded c4 record alone (bench.sharded_record)."""
import argparse, json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
args = argparse.Namespace(dist='room', pieces=2)
rec = bench.sharded_record(args, dist.get_world_size(), rank, dev)
if rank == 0:
  print(json.dumps({k: (round(v['ms_per_step'], 4) if isinstance(v, dict) and 'ms_per_step' in v else v) for k, v in rec.items()}))
dist.destroy_process_group()
