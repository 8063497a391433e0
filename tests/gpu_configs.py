"""Measure the BASELINE.json side configs on one B200 and the torch-eager bars
(BASELINE.md section 5): writes gpurun_out/configs_r1.json.
  C2 single-pair latency is in bench.py; here: C4 input pipeline (batch 512),
  C5 batched inference 480x640 / 720x1280 (batch 32), torch-eager fp32 / bf16-autocast
  train step and inference on the same GPU (oracle functional model = the reference's ops)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stereo_oracle as so
from stereo_depth_estimation_b200 import StereoUNet
from stereo_depth_estimation_b200.preprocess import AugmentSampler, DevicePreprocessor
from stereo_depth_estimation_b200.step import FusedStep

dev = torch.device("cuda:0")
out = {}

def timed(fn, warm=3, iters=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / iters

# ---- C4: input pipeline only, 540x960 -> 240x320, batch 512, augment on
B = 512
rng = np.random.default_rng(0)
src = [torch.from_numpy(rng.integers(0, 256, (B, 540, 960, 3), dtype=np.uint8)).to(dev) for _ in range(3)]
pre = DevicePreprocessor(dev, B, (240, 320))
sampler = AugmentSampler(seed=0)
buf = None
def run_pre():
    global buf
    buf = pre(src[0], src[1], src[2], aug=sampler.sample_packed(B), out=buf)
ms = timed(run_pre)
def run_pre_noaug():
    global buf
    buf = pre(src[0], src[1], src[2], out=buf)
ms_na = timed(run_pre_noaug)
bytes_per_sample = 6892800
out["C4_input_pipeline_b512"] = {"ms_per_batch_aug": ms, "samples_per_s_aug": B / ms * 1e3,
    "ms_per_batch_noaug": ms_na, "samples_per_s_noaug": B / ms_na * 1e3,
    "algorithmic_GBps_aug": B * bytes_per_sample / ms / 1e6, "algorithmic_GBps_noaug": B * bytes_per_sample / ms_na / 1e6}
del src, buf; pre.close(); torch.cuda.empty_cache()

# ---- C5: batched inference at scaled resolutions
torch.manual_seed(0)
model = StereoUNet().to(dev).eval()
flops = {(240, 320): 28.430e9, (480, 640): 113.718e9, (720, 1280): 341.154e9}
for (h, w, b) in [(240, 320, 32), (480, 640, 32), (720, 1280, 32)]:
    x = torch.rand(b, 6, h, w, device=dev)
    with torch.inference_mode():
        ms = timed(lambda: model(x, return_uncertainty=True), 3, 10)
    out[f"C5_infer_b{b}_{h}x{w}"] = {"ms_per_batch": ms, "pairs_per_s": b / ms * 1e3, "TFLOPs": b * flops[(h, w)] / ms / 1e9}
    del x
model._engine.close(); del model; torch.cuda.empty_cache()

# ---- torch-eager bars on the same GPU (cuDNN), fp32 (no TF32) and bf16 autocast
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
sd = {k: v.to(dev) for k, v in so.init_state_dict(42).items()}
def eager_train(b, autocast):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(b, 6, 240, 320, generator=g).to(dev); t = (torch.rand(b, 1, 240, 320, generator=g) * 2).to(dev)
    m = t > 0.2
    leaves = {k: sd[k].clone().requires_grad_(True) for k in so.param_keys(sd)}
    work = dict(sd); work.update(leaves)
    opt = torch.optim.AdamW(list(leaves.values()), lr=1e-3, weight_decay=1e-4)
    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            d, lv = so.model_forward(work, x, True, True, {})
        loss, _ = so.loss_and_sums(d.float(), lv.float(), t, m)
        loss.backward(); opt.step()
    return timed(step, 2, 5)
for b, ac in [(64, False), (64, True)]:
    ms = eager_train(b, ac)
    out[f"torch_eager_train_b{b}_{'bf16' if ac else 'fp32'}"] = {"ms_per_step": ms, "pairs_per_s": b / ms * 1e3}
def eager_infer(b, h, w, autocast):
    x = torch.rand(b, 6, h, w, device=dev)
    def f():
        with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            so.model_forward(sd, x, False, True)
    return timed(f, 5, 50 if b == 1 else 10)
for b, h, w in [(1, 240, 320), (32, 240, 320), (32, 480, 640)]:
    for ac in (False, True):
        ms = eager_infer(b, h, w, ac)
        out[f"torch_eager_infer_b{b}_{h}x{w}_{'bf16' if ac else 'fp32'}"] = {"ms": ms, "pairs_per_s": b / ms * 1e3}
# ours, train step at batch 64 for a like-for-like against the eager numbers
torch.manual_seed(0)
model = StereoUNet().to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
step = FusedStep(model, opt)
g = torch.Generator().manual_seed(1)
x = torch.rand(64, 6, 240, 320, generator=g).to(dev); t = (torch.rand(64, 1, 240, 320, generator=g) * 2).to(dev)
batch = {"input": x, "target": t, "valid_mask": t > 0.2}
ms = timed(lambda: step.train_step(batch), 3, 10)
out["ours_train_b64_preassembled_batch"] = {"ms_per_step": ms, "pairs_per_s": 64 / ms * 1e3}
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_r1.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
