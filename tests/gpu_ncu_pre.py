"""Small workload for ncu: the input pipeline alone (decode + resize + augment) on 64 raw 540x960 samples."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_estimation_b200.preprocess import AugmentSampler, DevicePreprocessor
dev = torch.device("cuda:0")
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0)
src = [torch.from_numpy(rng.integers(0, 256, (b, 540, 960, 3), dtype=np.uint8)).to(dev) for _ in range(3)]
pre = DevicePreprocessor(dev, b, (240, 320))
s = AugmentSampler(seed=0)
out = None
for _ in range(3):
    out = pre(src[0], src[1], src[2], aug=s.sample_packed(b), out=out)
torch.cuda.synchronize()
print("ok")
