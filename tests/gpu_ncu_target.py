"""Small workload for ncu: one train step (fwd + bwd) at batch 16, 240x320."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_estimation_b200 import StereoUNet
from stereo_depth_estimation_b200.step import FusedStep
dev = torch.device("cuda:0")
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
model = StereoUNet().to(dev)
g = torch.Generator().manual_seed(1)
x = torch.rand(b, 6, 240, 320, generator=g).to(dev)
t = (torch.rand(b, 1, 240, 320, generator=g) * 2).to(dev)
batch = {"input": x, "target": t, "valid_mask": t > 0.2}
step = FusedStep(model, None)
for _ in range(2):
    step.train_step(batch)
torch.cuda.synchronize()
print("ok")
