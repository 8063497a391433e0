"""Per-role clock64 trace of one wgrad launch (needs the -DSDN_FORENSICS build, SDN_DEBUG_TRACE_WGRAD=layer)."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_depth_estimation_b200 import StereoUNet, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = StereoUNet().to(dev)
x = torch.rand(64, 6, 240, 320, device=dev)
model.train()
for _ in range(2):
    d, lv = model(x, return_uncertainty=True)
    (d.mean() + lv.mean()).backward()
buf = np.zeros(3 * 16 * 8, dtype=np.int64)
_lib.check(_lib.load().sdn_debug_trace(model._engine.ctx, buf.ctypes.data))
t = buf.reshape(3, 16, 8)
t0 = t[t > 0].min()
for r, name in enumerate(["producer", "mma"]):
    print(name)
    for tile in range(2, 10):
        print("  tile", tile, " ".join(f"{(v - t0) if v > 0 else -1:7d}" for v in t[r, tile][:4]))
