"""Calibration (not a test): how far does torch's OWN bf16 autocast drift from its fp32
path on this network?  Gives the scale against which our bf16 tolerances are stated."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stereo_oracle as so
from tests.gpu_bringup import make_batch, rel
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
b, h, w = (int(v) for v in sys.argv[1:4])
sd = {k: v.to(dev) for k, v in so.init_state_dict(42).items()}
x, t, m = make_batch(b, h, w)
res = {}
for mode in ("fp32", "bf16", "fp32_again"):
    leaves = {k: sd[k].clone().requires_grad_(True) for k in so.param_keys(sd)}
    work = dict(sd); work.update(leaves)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        disp, logvar = so.model_forward(work, x, True, True, {})
    loss, _ = so.loss_and_sums(disp.float(), logvar.float(), t, m)
    loss.backward()
    res[mode] = (disp.detach().float(), logvar.detach().float(), loss.item(), {k: v.grad for k, v in leaves.items()})
for mode in ("bf16", "fp32_again"):
    d, l, ls, g = res[mode]; d0, l0, ls0, g0 = res["fp32"]
    print(mode, "disp rel", rel(d, d0)[0], "max/max", (d - d0).abs().max().item() / d0.abs().max().item(),
          "logvar rel", rel(l, l0)[0], "loss rel", (ls - ls0) / ls0)
    worst = 0
    for k in g0:
        r = rel(g[k], g0[k])[0]; worst = max(worst, r)
        cos = torch.nn.functional.cosine_similarity(g[k].flatten().double(), g0[k].flatten().double(), dim=0).item()
        if mode == "bf16": print(f"   {k:28s} rel={r:.3e} cos={cos:.4f}")
    print(mode, "worst grad rel", worst)
