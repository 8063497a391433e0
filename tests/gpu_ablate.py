"""Timing ablation of the forward conv kernels (debug flags, results are garbage)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_estimation_b200 import StereoUNet
dev = torch.device("cuda:0")
b = 64
torch.manual_seed(0)
model = StereoUNet().to(dev)
x = torch.rand(b, 6, 240, 320, device=dev)
model.train()
with torch.no_grad():
    for _ in range(2):
        model(x, return_uncertainty=True)
    model.profile_enable(True)
    for _ in range(3):
        model(x, return_uncertainty=True)
    rows = model.profile_dump()
names=[f"{bk}.{i}" for bk in ["enc1","enc2","enc3","enc4","bott","dec4","dec3","dec2","dec1"] for i in (0,3)]
print(os.environ.get("SDN_DEBUG_ABLATE","0"), ' '.join(f"{names[r['layer']]}:{r['ms']/3:.3f}" for r in sorted([r for r in rows if r['name']=='conv_fprop'], key=lambda r:r['layer'])))
