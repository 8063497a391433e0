"""Small end-to-end train step for compute-sanitizer (memcheck): preprocess -> fused step -> AdamW, twice,
plus an eval forward.  Shapes chosen so that tiles are ragged (H not a multiple of 16)."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_estimation_b200 import StereoUNet
from stereo_depth_estimation_b200.optim import FusedAdamW
from stereo_depth_estimation_b200.preprocess import AugmentSampler, DevicePreprocessor
from stereo_depth_estimation_b200.step import FusedStep

dev = torch.device("cuda:0")
B, H, W = 3, 48, 80
torch.manual_seed(0)
model = StereoUNet().to(dev)
opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
step = FusedStep(model, opt)
pre = DevicePreprocessor(dev, B, (H, W))
rng = np.random.default_rng(0)
src = [torch.from_numpy(rng.integers(0, 256, (B, 108, 240, 3), dtype=np.uint8)).to(dev) for _ in range(3)]
src[2][..., 0] = src[2][..., 0] % 4
count = torch.zeros(1, dtype=torch.int64, device=dev)
sampler = AugmentSampler(seed=0)
for _ in range(2):
    out = pre(src[0], src[1], src[2], aug=sampler.sample_packed(B), count_out=count)
    step.train_step(out, valid_count=count)
torch.cuda.synchronize()
model.eval()
with torch.inference_mode():
    d, lv = model(out["input"], return_uncertainty=True)
torch.cuda.synchronize()
# rows N3 / N4: cache-format entry and the live-view pre / post kernels
from stereo_depth_estimation_b200 import LivePipeline
l8 = torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).to(dev)
d16 = torch.from_numpy((rng.random((B, H, W)) * 50).astype(np.float16)).to(dev)
pre.from_cache(l8, l8, d16, aug=sampler.sample_packed(B), count_out=count)
live = LivePipeline(model, model_size=(W, H), ema_alpha=0.5, focal_length_px=100.0, baseline_m=0.07)
f0 = rng.integers(0, 256, (123, 157, 3), dtype=np.uint8)
maps = live(f0, f0)
maps = live(f0, f0)
torch.cuda.synchronize()
print("ok", float(d.mean()), step.read_metrics(), float(np.nanmean(maps["depth"])))
