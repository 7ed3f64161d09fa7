"""Profiling driver (run under ncu): a few train steps of the bench workload, nothing else."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendpd_b200 import models
from opendpd_b200.train import NativeTrainStep
kind = sys.argv[1] if len(sys.argv) > 1 else "dgru"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 13
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
T = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
torch.manual_seed(0)
net = models.CoreModel(2, H, 1, kind, num_dvr_units=3, thx=0.01, thh=0.05).cuda()
tr = NativeTrainStep(net)
x = (0.2 * torch.randn(B, T, 2)).cuda(); y = 0.9 * x
for _ in range(steps):
    tr.step(x, y)
torch.cuda.synchronize()
print("done")
