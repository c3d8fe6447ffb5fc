import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
class A: dist='room'
torch.cuda.set_device(0)
print(json.dumps(bench.compat_record(torch, A, torch.device('cuda', 0))))
