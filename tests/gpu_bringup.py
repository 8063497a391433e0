"""GPU bring-up diagnostics (not a pytest): per-layer error report of the CUDA path
against the oracle, printed so one gpurun call localises a bug.

  python tests/gpu_bringup.py pre|fwd|bwd [B H W]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import stereo_oracle as so  # noqa: E402
from stereo_depth_estimation_b200 import StereoUNet  # noqa: E402
from stereo_depth_estimation_b200.preprocess import DevicePreprocessor, ViewAug  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item(), (a - b).abs().max().item(), b.abs().max().item()


def ref_forward_capture(sd, x, training):
    """oracle model_forward with every pre-BN conv output and post-ReLU activation kept."""
    ys, acts, ups = [], [], []

    def block(x, name):
        for conv_i, bn_i in ((0, 1), (3, 4)):
            y = F.conv2d(x, sd[f"{name}.block.{conv_i}.weight"], None, padding=1)
            ys.append(y)
            x = F.relu(F.batch_norm(y, sd[f"{name}.block.{bn_i}.running_mean"].clone(),
                                    sd[f"{name}.block.{bn_i}.running_var"].clone(),
                                    sd[f"{name}.block.{bn_i}.weight"], sd[f"{name}.block.{bn_i}.bias"], training, 0.1, 1e-5))
            acts.append(x)
        return x

    s1 = block(x, "enc1")
    s2 = block(F.max_pool2d(s1, 2), "enc2")
    s3 = block(F.max_pool2d(s2, 2), "enc3")
    s4 = block(F.max_pool2d(s3, 2), "enc4")
    d = block(F.max_pool2d(s4, 2), "bottleneck")
    for level, skip in ((4, s4), (3, s3), (2, s2), (1, s1)):
        up = F.conv_transpose2d(d, sd[f"up{level}.weight"], sd[f"up{level}.bias"], stride=2)
        ups.append(up)
        d = block(torch.cat([up, skip], 1), f"dec{level}")
    disp = F.softplus(F.conv2d(d, sd["disparity_head.weight"], sd["disparity_head.bias"]))
    logvar = F.conv2d(d, sd["logvar_head.weight"], sd["logvar_head.bias"]).clamp(-6.0, 3.0)
    return ys, acts, ups, disp, logvar


def stage_pre():
    rng = np.random.default_rng(5)
    for (b, hs, ws, h, w) in [(2, 54, 96, 32, 48), (3, 50, 70, 16, 48), (2, 540, 960, 240, 320)]:
        L = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
        R = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
        D = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
        D[..., 0] = rng.integers(0, 4, (b, hs, ws))
        D[rng.random((b, hs, ws)) < 0.1] = 0
        pp = DevicePreprocessor(dev, b, (h, w))
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        for four in (False, True):
            out = pp(torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev), torch.from_numpy(D).to(dev),
                     fourterm=four, count_out=cnt)
            torch.cuda.synchronize()
            bad = 0
            for i in range(b):
                ref = so.make_sample(L[i], R[i], D[i], (h, w), formula="fourterm" if four else "separable")
                bad += int((out["input"][i].cpu().numpy() != ref["input"]).sum())
                bad += int((out["target"][i].cpu().numpy() != ref["target"]).sum())
                bad += int((out["valid_mask"][i].cpu().numpy() != ref["valid_mask"]).sum())
            print(f"pre {b}x{hs}x{ws}->{h}x{w} fourterm={four}: mismatching elements = {bad}; count={cnt.item()} "
                  f"ref_count={int(out['valid_mask'].sum().item())}")
        # augmentation (no noise) against the oracle
        views = [ViewAug(1.1, 0.85, 1.2, 0.05, 0.9, 0.0, 0.0, 1), ViewAug(0.8, 1.2, 0.8, -0.09, 1.2, 0.7, 0.0, 2)] * b
        out = pp(torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev), torch.from_numpy(D).to(dev), aug=views)
        torch.cuda.synchronize()
        worst = 0.0
        for i in range(b):
            aug = [dict(brightness=v.brightness, contrast=v.contrast, saturation=v.saturation, hue=v.hue, gamma=v.gamma,
                        blur_sigma=v.blur_sigma) for v in views[2 * i:2 * i + 2]]
            ref = so.make_sample(L[i], R[i], D[i], (h, w), aug=aug)
            worst = max(worst, float(np.abs(out["input"][i].cpu().numpy() - ref["input"]).max()))
        print(f"pre augment {b}x{hs}x{ws}->{h}x{w}: max abs err = {worst:.3e}")
        # noise statistics
        views = [ViewAug(noise_std=0.05, noise_seed=7 + k) for k in range(2 * b)]
        base = pp(torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev), torch.from_numpy(D).to(dev),
                  aug=[ViewAug() for _ in range(2 * b)])["input"].clone()
        noisy = pp(torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev), torch.from_numpy(D).to(dev), aug=views)["input"]
        inner = (base > 0.2) & (base < 0.8)
        dlt = (noisy - base)[inner]
        print(f"pre noise: mean={dlt.mean().item():.4e} std={dlt.std().item():.4e} (want 0, 0.05) n={dlt.numel()}")
        pp.close()


def make_batch(b, h, w, seed=123):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 6, h, w, generator=g)
    scale = float(os.environ.get("SDN_TARGET_SCALE", "64.0"))
    t = torch.rand(b, 1, h, w, generator=g) * scale
    t[:, :, : h // 4, : w // 4] = 0.0
    return x.to(dev), t.to(dev), (t > 0).to(dev)


NAMES = [f"{b}.{i}" for b in so.BLOCKS for i in (0, 3)]


def stage_fwd(b, h, w):
    torch.manual_seed(42)
    model = StereoUNet().to(dev)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, t, m = make_batch(b, h, w)
    for training in (True, False):
        model.train(training)
        with torch.no_grad():
            t0 = time.time()
            disp, logvar = model(x, return_uncertainty=True)
            torch.cuda.synchronize()
            print(f"forward training={training} ok in {time.time() - t0:.3f}s, launches={model.launch_count()}")
            ys, acts, ups, rdisp, rlogvar = ref_forward_capture(sd, x, training)
        for i in range(18):
            y = model.debug_activation(i, 0).to(dev).permute(0, 3, 1, 2)
            a = model.debug_activation(i, 1).to(dev).permute(0, 3, 1, 2)
            ry, ra = rel(y, ys[i]), rel(a, acts[i])
            print(f"  L{i:02d} {NAMES[i]:14s} y rel={ry[0]:.3e} max={ry[1]:.3e}/{ry[2]:.2e} | a rel={ra[0]:.3e} max={ra[1]:.3e}/{ra[2]:.2e}")
        for k in range(4):
            u = model.debug_activation(100 + k, 0).to(dev).permute(0, 3, 1, 2)
            ru = rel(u, ups[k])
            print(f"  up{4 - k} rel={ru[0]:.3e} max={ru[1]:.3e}/{ru[2]:.2e}")
        rd, rl = rel(disp, rdisp), rel(logvar, rlogvar)
        print(f"  disp rel={rd[0]:.3e} max={rd[1]:.3e}/{rd[2]:.2e} | logvar rel={rl[0]:.3e} max={rl[1]:.3e}/{rl[2]:.2e}")
        print(f"  max rel err (elementwise, |ref|>1e-2): disp={((disp - rdisp).abs() / rdisp.abs().clamp(min=1e-2)).max().item():.3e} "
              f"logvar={((logvar - rlogvar).abs() / rlogvar.abs().clamp(min=1e-2)).max().item():.3e}")
        if training:
            # running statistics after one training forward
            new = {}
            so.model_forward(sd, x, True, True, new)
            worst = 0.0
            for k, v in new.items():
                if "running" in k:
                    worst = max(worst, rel(model.state_dict()[k], v)[0])
            nbt = model.state_dict()["enc1.block.1.num_batches_tracked"].item()
            print(f"  BN running stats worst rel={worst:.3e}; num_batches_tracked={nbt}")
            # restore buffers so the eval pass of both sides sees the same state
            sd = {k: v.detach().clone() for k, v in model.state_dict().items()}


def stage_bwd(b, h, w):
    torch.manual_seed(42)
    model = StereoUNet().to(dev)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, t, m = make_batch(b, h, w)
    model.train()
    disp, logvar = model(x, return_uncertainty=True)
    loss, sums = so.loss_and_sums(disp, logvar, t, m)
    loss.backward()
    torch.cuda.synchronize()
    print(f"backward ok; loss={loss.item():.6f} launches={model.launch_count()}")
    # oracle
    leaves = {k: sd[k].clone().requires_grad_(True) for k in so.param_keys(sd)}
    work = dict(sd)
    work.update(leaves)
    rdisp, rlogvar = so.model_forward(work, x, True, True, {})
    rloss, _ = so.loss_and_sums(rdisp, rlogvar, t, m)
    rloss.backward()
    print(f"ref loss={rloss.item():.6f} rel diff={(loss.item() - rloss.item()) / rloss.item():.3e}")
    for name, p in model.named_parameters():
        r = rel(p.grad, leaves[name].grad)
        print(f"  grad {name:28s} rel={r[0]:.3e} max={r[1]:.3e}/{r[2]:.2e}")


def stage_bwdiso(b, h, w):
    """Isolate each backward kernel: feed torch the CUDA path's OWN inputs of that
    kernel (read back through sdn_debug_read) and compare outputs."""
    torch.manual_seed(42)
    model = StereoUNet().to(dev)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, t, m = make_batch(b, h, w)
    model.train()
    disp, logvar = model(x, return_uncertainty=True)
    loss, _ = so.loss_and_sums(disp, logvar, t, m)
    loss.backward()
    torch.cuda.synchronize()
    nchw = lambda which, kind: model.debug_activation(which, kind).to(dev).permute(0, 3, 1, 2).contiguous()
    y = [nchw(i, 0) for i in range(18)]
    a = [nchw(i, 1) for i in range(18)]
    dy = [nchw(i, 2) for i in range(18)]
    ga = [nchw(i, 3) for i in range(18)]
    u = [nchw(100 + k, 0) for k in range(4)]
    gu = [nchw(100 + k, 3) for k in range(4)]
    params = dict(model.named_parameters())
    wname = lambda i: f"{so.BLOCKS[i // 2]}.block.{0 if i % 2 == 0 else 3}.weight"
    bnname = lambda i: f"{so.BLOCKS[i // 2]}.block.{1 if i % 2 == 0 else 4}"

    def layer_input(i):
        if i == 0:
            return x
        if i % 2 == 1:
            return a[i - 1]
        if i <= 8:
            return F.max_pool2d(a[i - 1], 2)
        k = (i - 10) // 2
        return torch.cat([u[k], a[7 - 2 * k]], 1)

    for i in range(17, -1, -1):
        w_ = params[wname(i)]
        xin = layer_input(i)
        # weight gradient kernel
        wg_ref = torch.nn.grad.conv2d_weight(xin, w_.shape, dy[i], padding=1)
        r = rel(w_.grad, wg_ref)
        # BN backward (dy from y, ga[, pooled grad]) recomputed in torch from the path's own tensors
        gamma, beta = params[bnname(i) + ".weight"], params[bnname(i) + ".bias"]
        yy = y[i].clone().requires_grad_(True)
        z = F.batch_norm(yy, None, None, gamma, beta, True, 0.1, 1e-5)
        act = F.relu(z)
        if i in (1, 3, 5, 7):
            nxt = i + 1
            wn = params[wname(nxt)]
            gp_ref = torch.nn.grad.conv2d_input(F.max_pool2d(a[i], 2).shape, wn, dy[nxt], padding=1)
            pooled = F.max_pool2d(act, 2)
            (act * ga[i]).sum().backward(retain_graph=True, inputs=[yy])
            g1 = yy.grad.clone(); yy.grad = None
            (pooled * gp_ref).sum().backward(inputs=[yy])
            dy_ref = g1 + yy.grad
        else:
            (act * ga[i]).sum().backward(inputs=[yy])
            dy_ref = yy.grad
        rb = rel(dy[i], dy_ref)
        msg = f"  L{i:02d} {NAMES[i]:14s} wgrad rel={r[0]:.3e} | bn_bwd dy rel={rb[0]:.3e}"
        # data gradient kernel
        if i > 0:
            gin_ref = torch.nn.grad.conv2d_input(xin.shape, w_, dy[i], padding=1)
            if i % 2 == 1:
                rd = rel(ga[i - 1], gin_ref)
                msg += f" | dgrad rel={rd[0]:.3e}"
            elif i >= 10:
                k = (i - 10) // 2
                c = gu[k].shape[1]
                rd1, rd2 = rel(gu[k], gin_ref[:, :c]), rel(ga[7 - 2 * k], gin_ref[:, c:])
                msg += f" | dgrad up rel={rd1[0]:.3e} skip rel={rd2[0]:.3e}"
        print(msg)
    for k in range(4):
        lvl = 4 - k
        wt, bt = params[f"up{lvl}.weight"], params[f"up{lvl}.bias"]
        src = a[9 + 2 * k].clone().requires_grad_(True)
        wl = wt.detach().clone().requires_grad_(True)
        bl = bt.detach().clone().requires_grad_(True)
        out = F.conv_transpose2d(src, wl, bl, stride=2)
        (out * gu[k]).sum().backward()
        print(f"  up{lvl}: wgrad rel={rel(wt.grad, wl.grad)[0]:.3e} bias rel={rel(bt.grad, bl.grad)[0]:.3e} "
              f"dgrad rel={rel(ga[9 + 2 * k], src.grad)[0]:.3e}")


if __name__ == "__main__":
    stage = sys.argv[1]
    b, h, w = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (2, 32, 48)
    print(f"== {stage} B={b} H={h} W={w} device={torch.cuda.get_device_name(0)}")
    if stage == "pre":
        stage_pre()
    elif stage == "fwd":
        stage_fwd(b, h, w)
    elif stage == "bwd":
        stage_bwd(b, h, w)
    elif stage == "bwdiso":
        stage_bwdiso(b, h, w)
    print("== done")
