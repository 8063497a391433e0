"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI via the Python
mirror, against the oracle and the committed golden fixtures.

Tolerances (stated here, see DESIGN.md "Parity"):
  * decode / resize / mask / valid count : bit-exact
  * augmentation (explicit parameters)   : <= 1e-5 abs (transcendentals differ by ulps)
  * eval-mode forward (live-view path)   : max|err| / max|ref| <= 1e-2  (north_star)
  * train-mode forward, bf16 storage     : rel-L2 <= 1.5e-2, loss <= 1e-3 relative
  * every backward kernel on ITS OWN inputs vs torch fp32: wgrad <= 1e-4, dgrad /
    BatchNorm-backward <= 6e-3 rel-L2 (one bf16 rounding of the output)
  * end-to-end gradients vs fp32 autograd: no worse than 1.5x torch's own bf16-autocast
    drift on the same batch (+2e-2), measured in the same test
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import stereo_oracle as so

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dev():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()


def make_batch(dev, b, h, w, seed=123, scale=1.5):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 6, h, w, generator=g)
    t = torch.rand(b, 1, h, w, generator=g) * scale
    t[:, :, : h // 4, : w // 4] = 0.0
    return {"input": x.to(dev), "target": t.to(dev), "valid_mask": (t > 0).to(dev)}


def fresh_model(dev, seed=42):
    from stereo_depth_estimation_b200 import StereoUNet

    torch.manual_seed(seed)
    model = StereoUNet().to(dev)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return model, sd


def synth_sources(rng, b, hs, ws):
    L = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
    R = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
    D = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
    D[..., 0] = rng.integers(0, 4, (b, hs, ws))
    D[rng.random((b, hs, ws)) < 0.1] = 0
    return L, R, D


# ------------------------------------------------------------------ preprocess
@pytest.mark.parametrize("name", ["sample_exact", "sample_ragged", "sample_up", "sample_same"])
def test_preprocess_matches_reference_fixture_bit_exact(dev, name):
    """The fixtures are outputs of the real reference (small images -> its 'fourterm' CPU kernel)."""
    from stereo_depth_estimation_b200.preprocess import DevicePreprocessor

    f = np.load(os.path.join(GOLDEN, name + ".npz"))
    h, w = (int(v) for v in f["out_hw"])
    pre = DevicePreprocessor(dev, 1, (h, w))
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    args = [torch.from_numpy(f[k][None]).to(dev) for k in ("left", "right", "disp")]
    out = pre(*args, fourterm=True, count_out=cnt)
    assert np.array_equal(out["input"][0].cpu().numpy(), f["input"])
    assert np.array_equal(out["target"][0].cpu().numpy(), f["target"])
    assert np.array_equal(out["valid_mask"][0].cpu().numpy(), f["valid_mask"])
    assert out["valid_mask"].dtype == torch.bool and out["input"].dtype == torch.float32
    assert int(cnt.item()) == int(f["valid_mask"].sum())
    sep = pre(*args)  # canonical ordering == oracle default
    ref = so.make_sample(f["left"], f["right"], f["disp"], (h, w))
    assert np.array_equal(sep["input"][0].cpu().numpy(), ref["input"])
    assert np.array_equal(sep["target"][0].cpu().numpy(), ref["target"])


def test_preprocess_full_size_bit_exact(dev):
    """BASELINE.json config 4 shape (540x960 -> 240x320), ragged batch of 3."""
    from stereo_depth_estimation_b200.preprocess import DevicePreprocessor

    rng = np.random.default_rng(5)
    L, R, D = synth_sources(rng, 3, 540, 960)
    D[1] = 0  # a sample with no valid pixel at all
    pre = DevicePreprocessor(dev, 4, (240, 320))
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    out = pre(torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev), torch.from_numpy(D).to(dev), count_out=cnt)
    total = 0
    for i in range(3):
        ref = so.make_sample(L[i], R[i], D[i], (240, 320))
        assert np.array_equal(out["input"][i].cpu().numpy(), ref["input"])
        assert np.array_equal(out["target"][i].cpu().numpy(), ref["target"])
        assert np.array_equal(out["valid_mask"][i].cpu().numpy(), ref["valid_mask"])
        total += int(ref["valid_mask"].sum())
    assert int(cnt.item()) == total
    assert not out["valid_mask"][1].any()
    # horizontal scale 3.0 is pure point sampling of column 3x+1
    plain = (L[0, :, 1::3, :].astype(np.float32) / np.float32(255.0))
    assert out["input"].shape == (3, 6, 240, 320) and plain.shape[1] == 320


def test_preprocess_decode_kats(dev):
    """reference tests/test_dataset.py:31-35,38-61 through the device kernel."""
    from stereo_depth_estimation_b200.preprocess import DevicePreprocessor

    src = np.full((1, 2, 4), 1.5, dtype=np.float32)
    enc = so.encode_disparity(src)
    zeros = np.zeros_like(enc)
    pre = DevicePreprocessor(dev, 1, (16, 32))  # the library needs multiples of 16
    big = np.repeat(np.repeat(enc, 8, axis=1), 4, axis=2)  # 16 x 16 constant 1.5 px
    out = pre(torch.from_numpy(np.zeros_like(big)).to(dev), torch.from_numpy(np.zeros_like(big)).to(dev),
              torch.from_numpy(big).to(dev))
    np.testing.assert_allclose(out["target"].cpu().numpy(), 3.0, atol=1e-3)  # width doubled -> disparity doubled
    white = np.full((1, 16, 32, 3), 255, dtype=np.uint8)
    out = pre(torch.from_numpy(white).to(dev), torch.from_numpy(white).to(dev), torch.from_numpy(white).to(dev))
    assert np.all(out["target"].cpu().numpy() == np.float32(16646.654))
    assert np.all(out["input"].cpu().numpy() == np.float32(1.0))
    _ = zeros


def test_augment_matches_oracle(dev):
    from stereo_depth_estimation_b200.preprocess import DevicePreprocessor, ViewAug

    rng = np.random.default_rng(9)
    b, h, w = 3, 48, 64
    L, R, D = synth_sources(rng, b, 108, 192)
    L[0, :20, :40] = 128  # gray patch: maxc == minc branch
    views = [ViewAug(1.1, 0.85, 1.2, 0.05, 0.9), ViewAug(0.8, 1.2, 0.8, -0.09, 1.2, blur_sigma=0.7),
             ViewAug(), ViewAug(1.2, 0.8, 1.25, 0.5, 0.8, blur_sigma=1.0),
             ViewAug(0.9, 1.1, 0.0, -0.5, 1.0), ViewAug(1.0, 1.0, 1.0, 0.0, 1.0, blur_sigma=0.1)]
    pre = DevicePreprocessor(dev, b, (h, w))
    out = pre(torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev), torch.from_numpy(D).to(dev), aug=views)
    for i in range(b):
        aug = [dict(brightness=v.brightness, contrast=v.contrast, saturation=v.saturation, hue=v.hue, gamma=v.gamma,
                    blur_sigma=v.blur_sigma) for v in views[2 * i: 2 * i + 2]]
        ref = so.make_sample(L[i], R[i], D[i], (h, w), aug=aug)
        np.testing.assert_allclose(out["input"][i].cpu().numpy(), ref["input"], atol=1e-5, rtol=0)
        assert np.array_equal(out["target"][i].cpu().numpy(), ref["target"])  # augmentation never touches the target


def test_augment_noise_statistics(dev):
    from stereo_depth_estimation_b200.preprocess import DevicePreprocessor, ViewAug

    rng = np.random.default_rng(2)
    b, h, w = 2, 240, 320
    L, R, D = synth_sources(rng, b, 240, 320)
    pre = DevicePreprocessor(dev, b, (h, w))
    src = [torch.from_numpy(a).to(dev) for a in (L, R, D)]
    base = pre(*src, aug=[ViewAug() for _ in range(2 * b)])["input"].clone()
    noisy = pre(*src, aug=[ViewAug(noise_std=0.05, noise_seed=10 + k) for k in range(2 * b)])["input"]
    inner = (base > 0.25) & (base < 0.75)
    d = (noisy - base)[inner]
    assert abs(d.mean().item()) < 5e-4 and abs(d.std().item() - 0.05) < 5e-4
    assert noisy.min().item() >= 0.0 and noisy.max().item() <= 1.0
    again = pre(*src, aug=[ViewAug(noise_std=0.05, noise_seed=10 + k) for k in range(2 * b)])["input"]
    assert torch.equal(noisy, again)  # counter-based generator: same seed, same noise
    v0, v1 = (noisy - base)[0, :3][inner[0, :3]], (noisy - base)[0, 3:][inner[0, 3:]]
    n = min(v0.numel(), v1.numel())
    assert abs(torch.corrcoef(torch.stack([v0[:n], v1[:n]]))[0, 1].item()) < 0.02  # L / R independent


# --------------------------------------------------------------------- forward
@pytest.mark.parametrize("b,h,w", [(1, 240, 320), (3, 64, 96), (2, 32, 48)])
def test_forward_eval_matches_oracle(dev, b, h, w):
    model, sd = fresh_model(dev)
    batch = make_batch(dev, b, h, w)
    model.eval()
    with torch.inference_mode():
        disp, logvar = model(batch["input"], return_uncertainty=True)
        disp_only = model(batch["input"])
    rd, rl = so.model_forward(sd, batch["input"], False, True)
    assert disp.shape == (b, 1, h, w) and disp.dtype == torch.float32
    assert torch.equal(disp, disp_only)
    assert (disp - rd).abs().max().item() / rd.abs().max().item() <= 1e-2
    assert (logvar - rl).abs().max().item() / rl.abs().max().item() <= 1e-2
    assert disp.min().item() >= 0.0 and logvar.min().item() >= -6.0 and logvar.max().item() <= 3.0


def test_forward_eval_graph_replay_matches_oracle(dev):
    """The live viewer's per-frame call (depth_live_dl.py:518-529) replays a CUDA graph from the SECOND call
    per (batch, want_logvar) on; the replayed outputs - not only the eager first call - must match."""
    model, sd = fresh_model(dev)
    model.eval()
    for want in (True, False):
        for call in range(3):
            x = make_batch(dev, 1, 240, 320, seed=500 + call)["input"]
            with torch.inference_mode():
                out = model(x, return_uncertainty=want)
            if want:
                assert (1, True) in model._engine.graphs
            rd, rl = so.model_forward(sd, x, False, True)
            disp, logvar = out if want else (out, None)
            assert (disp - rd).abs().max().item() / rd.abs().max().item() <= 1e-2, (want, call)
            if want:
                assert (logvar - rl).abs().max().item() / rl.abs().max().item() <= 1e-2, (want, call)
    # a weight update invalidates the captured graphs (the operand cache is re-packed)
    with torch.no_grad():
        model.disparity_head.bias.add_(1.0)
    x = make_batch(dev, 1, 240, 320, seed=510)["input"]
    sd2 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for call in range(2):
        with torch.inference_mode():
            disp = model(x)
        rd, _ = so.model_forward(sd2, x, False, True)
        assert (disp - rd).abs().max().item() / rd.abs().max().item() <= 1e-2, call


@pytest.mark.parametrize("b,h,w", [(2, 480, 640), (1, 720, 1280), (32, 240, 320)])
def test_forward_eval_scaled_resolutions(dev, b, h, w):
    """BASELINE.json config 5 shapes (480x640, 720x1280) and a batch beyond the CUDA-graph limit."""
    model, sd = fresh_model(dev)
    x = make_batch(dev, b, h, w, seed=77)["input"]
    model.eval()
    with torch.inference_mode():
        disp, logvar = model(x, return_uncertainty=True)
    rd, rl = so.model_forward(sd, x, False, True)
    e_d = (disp - rd).abs().max().item() / rd.abs().max().item()
    e_l = (logvar - rl).abs().max().item() / rl.abs().max().item()
    print(f"eval {b}x{h}x{w}: disparity max-rel {e_d:.2e} logvar max-rel {e_l:.2e}")
    assert e_d <= 1e-2 and e_l <= 1e-2


@pytest.mark.parametrize("b", [17, 19])
def test_cta_pairs_odd_image_count(dev, b):
    """The deep layers run as CTA pairs (tcgen05 cta_group::2): the two CTAs of a cluster take consecutive image
    groups of one tile position, so with an ODD group count the last pair's second CTA works on an image past the
    batch (TMA zero-fills its loads and clips its stores, its rows are excluded from the BatchNorm statistics).
    Eval and train forward plus the train-mode BatchNorm buffers vs the fp32 oracle at such batches."""
    h, w = 64, 96
    model, sd = fresh_model(dev)
    batch = make_batch(dev, b, h, w, seed=300 + b)
    model.eval()
    with torch.inference_mode():
        disp, logvar = model(batch["input"], return_uncertainty=True)
    rd, rl = so.model_forward(sd, batch["input"], False, True)
    assert (disp - rd).abs().max().item() / rd.abs().max().item() <= 1e-2
    assert (logvar - rl).abs().max().item() / rl.abs().max().item() <= 1e-2
    model.train()
    with torch.no_grad():
        disp, logvar = model(batch["input"], return_uncertainty=True)
    bufs = {}
    rd, rl = so.model_forward(sd, batch["input"], True, True, bufs)
    e_d, e_l = rel(disp, rd), rel(logvar, rl)
    print(f"train {b}x{h}x{w} (odd pair count) vs fp32 oracle, rel-L2: disparity {e_d:.2e} logvar {e_l:.2e}")
    assert e_d <= 1e-2 and e_l <= 1.5e-2
    after = model.state_dict()
    for k, v in bufs.items():          # running statistics of every layer (the deep ones come from pair kernels)
        if "running" in k:
            assert rel(after[k], v.to(dev)) <= 2e-2, k


def test_forward_train_batch32_full_resolution(dev):
    """The per-GPU batch of the 8-GPU run (32 x 240 x 320), train mode, vs the fp32 oracle on the same GPU."""
    model, sd = fresh_model(dev)
    batch = make_batch(dev, 32, 240, 320, seed=78)
    model.train()
    with torch.no_grad():
        disp, logvar = model(batch["input"], return_uncertainty=True)
    rd, rl = so.model_forward(sd, batch["input"], True, True, {})
    with torch.autocast("cuda", dtype=torch.bfloat16):      # the yardstick: torch's own bf16 path, same weights
        td, tl = so.model_forward(sd, batch["input"], True, True, {})
    td, tl = td.float(), tl.float()
    loss, _ = so.loss_and_sums(disp, logvar, batch["target"], batch["valid_mask"])
    rloss, _ = so.loss_and_sums(rd, rl, batch["target"], batch["valid_mask"])
    e_loss = abs(loss.item() - rloss.item()) / abs(rloss.item())
    e_d, e_l, t_d, t_l = rel(disp, rd), rel(logvar, rl), rel(td, rd), rel(tl, rl)
    print(f"train 32x240x320 vs fp32 oracle, rel-L2: disparity ours {e_d:.2e} / torch-autocast {t_d:.2e}; "
          f"logvar ours {e_l:.2e} / torch-autocast {t_l:.2e}; loss rel {e_loss:.2e}")
    # north_star: <= 1e-2 on disparity / logvar, loss <= 1e-3.  Disparity and the loss meet it.  Train-mode logvar
    # at random init sits at ~1.2e-2 for ANY bf16-storage implementation (torch autocast measured beside it):
    # BatchNorm's mean subtraction amplifies the 2^-9 rounding of the stored conv outputs.  Stated deviation
    # (DESIGN section 4): logvar <= 1.5e-2 and no worse than 1.25x torch-autocast's own error.
    assert e_d <= 1e-2 and e_loss <= 1e-3
    assert e_l <= 1.5e-2 and e_l <= 1.25 * t_l + 1e-3


def test_eval_step_same_weights_matches_oracle(dev):
    """Row N2: validation path (run_epoch with optimizer=None, train.py:618) - eval-mode forward + the five
    metric sums - against the oracle at IDENTICAL weights and non-trivial BatchNorm buffers."""
    from stereo_depth_estimation_b200.step import FusedStep, run_epoch

    model, sd = fresh_model(dev)
    batches = [make_batch(dev, 3, 64, 96, seed=900 + i) for i in range(3)]
    model.train()
    with torch.no_grad():
        for b_ in batches:
            model(b_["input"])                      # three BatchNorm buffer updates
    cur = {k: v.detach().clone() for k, v in model.state_dict().items()}
    step = FusedStep(model, optimizer=None)
    step._ensure(dev)
    step.reset_metrics()
    tot = {"nll": 0.0, "abs": 0.0, "sq": 0.0, "sigma": 0.0, "count": 0}
    for b_ in batches:
        disp, logvar = step.eval_step(b_, want_outputs=True)
        rd, rl = so.model_forward(cur, b_["input"], False, True)
        assert (disp - rd).abs().max().item() / rd.abs().max().item() <= 1e-2
        assert (logvar - rl).abs().max().item() / rl.abs().max().item() <= 1e-2
        _, sums = so.loss_and_sums(rd, rl, b_["target"], b_["valid_mask"])
        for k in tot:
            tot[k] += sums[k]
    got = step.read_metrics()
    assert got["count"] == tot["count"]
    for k in ("nll", "abs", "sq", "sigma"):
        assert got[k] == pytest.approx(tot[k], rel=2e-3), k
    val, _ = run_epoch(model, batches, dev, optimizer=None)
    oval = so.run_epoch(cur, batches, None)
    for k in ("loss", "mae", "rmse", "sigma"):
        assert val[k] == pytest.approx(oval[k], rel=2e-3), k
    assert torch.equal(model.state_dict()["enc1.block.1.running_mean"], cur["enc1.block.1.running_mean"])


def test_forward_eval_golden_fixture(dev):
    """model.npz was produced by the real reference (seed 42, trained-one-step BN buffers)."""
    f = np.load(os.path.join(GOLDEN, "model.npz"))
    model, _ = fresh_model(dev)
    x = torch.from_numpy(f["input"]).to(dev)
    model.train()
    with torch.no_grad():
        d_t, l_t = model(x, return_uncertainty=True)   # updates the BN buffers like the reference run did
    assert rel(d_t.cpu(), torch.from_numpy(f["disp_train"])) < 2.5e-2   # B=2, 32x48: 12 samples per channel in the bottleneck
    model.eval()
    with torch.inference_mode():
        d_e, l_e = model(x, return_uncertainty=True)
    assert rel(d_e.cpu(), torch.from_numpy(f["disp_eval"])) < 2.5e-2
    assert rel(l_e.cpu(), torch.from_numpy(f["logvar_eval"])) < 5e-2
    for k, v in model.state_dict().items():
        if "num_batches" in k:
            assert int(v.item()) == 1


@pytest.mark.parametrize("b,h,w", [(4, 240, 320), (3, 64, 96)])
def test_forward_train_matches_oracle(dev, b, h, w):
    model, sd = fresh_model(dev)
    batch = make_batch(dev, b, h, w)
    model.train()
    with torch.no_grad():
        disp, logvar = model(batch["input"], return_uncertainty=True)
    new = {}
    rd, rl = so.model_forward(sd, batch["input"], True, True, new)
    e_d, e_l = rel(disp, rd), rel(logvar, rl)
    print(f"train {b}x{h}x{w} vs fp32 oracle, rel-L2: disparity {e_d:.2e} logvar {e_l:.2e}")
    # 4x240x320: north_star's 1e-2 on disparity; logvar as in test_forward_train_batch32_full_resolution.
    # 3x64x96 has only 72 samples per channel in the bottleneck's batch statistics: noisier, looser.
    assert e_d <= (1e-2 if h == 240 else 1.5e-2) and e_l <= (1.5e-2 if h == 240 else 3e-2)
    loss, _ = so.loss_and_sums(disp, logvar, batch["target"], batch["valid_mask"])
    rloss, _ = so.loss_and_sums(rd, rl, batch["target"], batch["valid_mask"])
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3
    cur = model.state_dict()
    for k, v in new.items():
        if "running" in k:
            assert rel(cur[k], v) < 2e-2, k
        elif "num_batches" in k:
            assert int(cur[k].item()) == int(v.item())


# -------------------------------------------------------------------- backward
def _nchw(model, which, kind, dev):
    return model.debug_activation(which, kind).to(dev).permute(0, 3, 1, 2).contiguous()


def test_backward_kernels_isolated(dev):
    """Each backward kernel against torch fp32 on the kernel's own inputs."""
    b, h, w = 3, 64, 96
    model, _ = fresh_model(dev)
    batch = make_batch(dev, b, h, w)
    model.train()
    disp, logvar = model(batch["input"], return_uncertainty=True)
    loss, _ = so.loss_and_sums(disp, logvar, batch["target"], batch["valid_mask"])
    loss.backward()
    y = [_nchw(model, i, 0, dev) for i in range(18)]
    a = [_nchw(model, i, 1, dev) for i in range(18)]
    dy = [_nchw(model, i, 2, dev) for i in range(18)]
    ga = [_nchw(model, i, 3, dev) for i in range(18)]
    u = [_nchw(model, 100 + k, 0, dev) for k in range(4)]
    gu = [_nchw(model, 100 + k, 3, dev) for k in range(4)]
    params = dict(model.named_parameters())
    wname = lambda i: f"{so.BLOCKS[i // 2]}.block.{0 if i % 2 == 0 else 3}.weight"  # noqa: E731
    bnname = lambda i: f"{so.BLOCKS[i // 2]}.block.{1 if i % 2 == 0 else 4}"  # noqa: E731

    def layer_input(i):
        if i == 0:
            return batch["input"]
        if i % 2 == 1:
            return a[i - 1]
        if i <= 8:
            return F.max_pool2d(a[i - 1], 2)
        k = (i - 10) // 2
        return torch.cat([u[k], a[7 - 2 * k]], 1)

    for i in range(18):
        wt = params[wname(i)]
        xin = layer_input(i)
        wg_ref = torch.nn.grad.conv2d_weight(xin, wt.shape, dy[i], padding=1)
        assert rel(wt.grad, wg_ref) < (6e-3 if i == 0 else 1e-4), f"wgrad {i}"  # layer 0 rounds its fp32 input to bf16
        if i > 0:
            gin = torch.nn.grad.conv2d_input(xin.shape, wt, dy[i], padding=1)
            if i % 2 == 1:
                assert rel(ga[i - 1], gin) < 6e-3, f"dgrad {i}"
            elif i >= 10:
                k = (i - 10) // 2
                c = gu[k].shape[1]
                assert rel(gu[k], gin[:, :c]) < 6e-3 and rel(ga[7 - 2 * k], gin[:, c:]) < 6e-3, f"dgrad {i}"
        yy = y[i].clone().requires_grad_(True)
        z = F.batch_norm(yy, None, None, params[bnname(i) + ".weight"], params[bnname(i) + ".bias"], True, 0.1, 1e-5)
        g_total = ga[i]
        if i in (1, 3, 5, 7):
            # pooled block outputs (model.py:83-86): the gradient of the pooled tensor is routed to the FIRST
            # maximum of each 2x2 quad of the bf16 activation the consumers saw (nn.MaxPool2d backward); torch's
            # own max_pool2d on that same activation gives the routing indices, ties included
            gp = _nchw(model, i, 4, dev)
            pooled, idx = F.max_pool2d(a[i], 2, return_indices=True)
            assert torch.equal(pooled, _nchw(model, i, 5, dev)), f"maxpool forward {i}"
            g_total = ga[i] + F.max_unpool2d(gp, idx, 2, output_size=a[i].shape[-2:])
        (F.relu(z) * g_total).sum().backward()
        assert rel(dy[i], yy.grad) < 6e-3, f"bn backward {i}"
    for k in range(4):
        lvl = 4 - k
        wt, bt = params[f"up{lvl}.weight"], params[f"up{lvl}.bias"]
        src = a[9 + 2 * k].clone().requires_grad_(True)
        wl, bl = wt.detach().clone().requires_grad_(True), bt.detach().clone().requires_grad_(True)
        (F.conv_transpose2d(src, wl, bl, stride=2) * gu[k]).sum().backward()
        assert rel(wt.grad, wl.grad) < 1e-4 and rel(bt.grad, bl.grad) < 1e-4
        assert rel(ga[9 + 2 * k], src.grad) < 6e-3


@pytest.mark.parametrize("path", ["module", "fused_step"])
def test_gradients_no_worse_than_torch_bf16_autocast(dev, path):
    """End-to-end gradients of BOTH entry points (autograd module path; FusedStep = sdn_train_step with the
    in-kernel loss seed) against the fp32 ORACLE, with torch's own bf16-autocast drift as the yardstick."""
    b, h, w = 4, 128, 160
    model, sd = fresh_model(dev)
    batch = make_batch(dev, b, h, w)
    model.train()
    if path == "module":
        disp, logvar = model(batch["input"], return_uncertainty=True)
        loss, _ = so.loss_and_sums(disp, logvar, batch["target"], batch["valid_mask"])
        loss.backward()
    else:
        from stereo_depth_estimation_b200.step import FusedStep

        assert FusedStep(model, optimizer=None).train_step(batch) > 0
    ours = {k: p.grad.clone() for k, p in model.named_parameters()}
    grads = {}
    for mode in ("fp32", "bf16"):
        leaves = {k: sd[k].clone().requires_grad_(True) for k in so.param_keys(sd)}
        work = dict(sd)
        work.update(leaves)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            rd, rl = so.model_forward(work, batch["input"], True, True, {})
        rloss, _ = so.loss_and_sums(rd.float(), rl.float(), batch["target"], batch["valid_mask"])
        rloss.backward()
        grads[mode] = {k: v.grad for k, v in leaves.items()}
    worst = (0.0, 0.0, "")
    for k in ours:
        e_ours = rel(ours[k], grads["fp32"][k])
        e_torch = rel(grads["bf16"][k], grads["fp32"][k])
        worst = max(worst, (e_ours, e_torch, k))
        assert e_ours <= 1.5 * e_torch + 2e-2, (k, e_ours, e_torch)
        cos = F.cosine_similarity(ours[k].flatten().double(), grads["fp32"][k].flatten().double(), dim=0).item()
        assert cos > 0.8, (k, cos)
    print(f"[{path}] worst gradient rel-L2 vs fp32 oracle: ours {worst[0]:.3e} torch-bf16-autocast {worst[1]:.3e} ({worst[2]})")
    for k in ("disparity_head.weight", "logvar_head.weight", "dec1.block.4.weight", "dec1.block.4.bias"):
        assert rel(ours[k], grads["fp32"][k]) < 2e-2, k


def test_autograd_contract_matches_reference_semantics(dev):
    """(1) a backward through a graph whose activations were overwritten by a later forward raises instead of
    returning gradients of the newer input; (2) forward(x) without the uncertainty head leaves
    logvar_head.{weight,bias}.grad None, like reference autograd (so AdamW skips them)."""
    model, _ = fresh_model(dev)
    model.train()
    b1, b2 = make_batch(dev, 2, 32, 48, seed=1), make_batch(dev, 2, 32, 48, seed=2)
    d1 = model(b1["input"])
    d2 = model(b2["input"])
    with pytest.raises(RuntimeError, match="ONE training forward"):
        d1.sum().backward()
    d2.sum().backward()
    assert model.logvar_head.weight.grad is None and model.logvar_head.bias.grad is None
    assert model.disparity_head.weight.grad is not None and model.enc1.block[0].weight.grad is not None
    # gradient accumulation over several forward/backward pairs is what autograd users do: still fine
    g1 = model.disparity_head.weight.grad.clone()
    model(b2["input"]).sum().backward()
    assert torch.allclose(model.disparity_head.weight.grad, 2 * g1, rtol=1e-3, atol=1e-6)


# ------------------------------------------------------------------ fused step
def test_fused_step_matches_module_path_and_oracle_sums(dev):
    from stereo_depth_estimation_b200.step import FusedStep

    b, h, w = 3, 64, 96
    batch = make_batch(dev, b, h, w)
    model_a, sd = fresh_model(dev)
    model_a.train()
    d, lv = model_a(batch["input"], return_uncertainty=True)
    loss, sums = so.loss_and_sums(d, lv, batch["target"], batch["valid_mask"])
    loss.backward()
    model_b, _ = fresh_model(dev)
    step = FusedStep(model_b, optimizer=None)
    n = step.train_step(batch)
    got = step.read_metrics()
    assert n == sums["count"] == got["count"]
    for k in ("nll", "abs", "sq", "sigma"):
        assert got[k] == pytest.approx(sums[k], rel=2e-4), k
    for (ka, pa), (kb, pb) in zip(model_a.named_parameters(), model_b.named_parameters()):
        # same kernels; only the loss seed differs (fused in-kernel vs torch ops + fp32 round trip) and the
        # wgrad atomics are unordered; BatchNorm's mean subtraction amplifies that in the deepest layers
        assert rel(pb.grad, pa.grad) < 3e-2, ka


def test_fused_step_skips_batch_without_valid_pixels(dev):
    from stereo_depth_estimation_b200.step import FusedStep

    model, sd = fresh_model(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    batch = make_batch(dev, 2, 32, 48)
    batch["valid_mask"] = torch.zeros_like(batch["valid_mask"])
    step = FusedStep(model, opt)
    assert step.train_step(batch) == 0
    for k, v in model.state_dict().items():
        if so.is_param_key(k):
            assert torch.equal(v, sd[k]), k        # no optimizer step (train.py:331-332)
    assert step.read_metrics()["count"] == 0
    nan_batch = make_batch(dev, 2, 32, 48)
    nan_batch["target"][0, 0, 20, 20] = float("nan")
    n = step.train_step(nan_batch)
    assert n == int((nan_batch["valid_mask"] & torch.isfinite(nan_batch["target"])).sum().item())
    assert all(torch.isfinite(p).all() for p in model.parameters())


def test_run_epoch_tracks_oracle_curve(dev):
    """A few-hundred-step synthetic training curve: fused CUDA steps (bf16 tensor cores)
    vs the fp32 oracle on the same data and the same AdamW settings.  The two
    trajectories are chaotic relative to each other at the 1e-3 level, so the curve is
    compared per 20-step segment with an 8 % (+0.01 absolute) band."""
    from stereo_depth_estimation_b200.step import run_epoch

    b, h, w, steps, seg = 4, 64, 96, 240, 20
    batches = [make_batch(dev, b, h, w, seed=300 + i) for i in range(24)]   # cycled
    model, sd = fresh_model(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    logged, ours, ref = [], [], []
    osd = {k: v.clone() for k, v in sd.items()}
    oopt = so.AdamWState()
    gs = 0
    for i in range(0, steps, seg):
        chunk = [batches[(i + j) % len(batches)] for j in range(seg)]
        m, gs = run_epoch(model, chunk, dev, optimizer=opt, global_step=gs, log_every_batches=10,
                          log_metrics=lambda d, step: logged.append((step, d)))
        ours.append(m["loss"])
        ref.append(so.run_epoch(osd, chunk, oopt)["loss"])
    assert gs == steps and len(logged) == steps // 10
    assert set(logged[0][1]) == {"train_loss_step", "train_nll_step", "train_mae_step", "train_rmse_step", "train_sigma_step"}
    for a_, r_ in zip(ours, ref):
        # the NLL crosses zero late in the run: relative band plus a small absolute floor
        assert abs(a_ - r_) <= 8e-2 * abs(r_) + 1e-2, (ours, ref)
    assert ours[-1] < ours[0] - 0.2 and ref[-1] < ref[0] - 0.2
    val, _ = run_epoch(model, batches[:2], dev, optimizer=None)
    oval = so.run_epoch(osd, batches[:2], None)
    assert val["mae"] == pytest.approx(oval["mae"], rel=1.5e-1)


def test_fused_adamw_matches_torch(dev):
    from stereo_depth_estimation_b200.optim import FusedAdamW

    torch.manual_seed(3)
    shapes = [(32, 6, 3, 3), (32,), (64, 32, 3, 3), (1,), (512, 256, 2, 2)]
    pa = [torch.randn(s, device=dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = torch.optim.AdamW(pa, lr=1e-3, weight_decay=1e-4)
    ob = FusedAdamW(pb, lr=1e-3, weight_decay=1e-4)
    gate = torch.ones(1, dtype=torch.int64, device=dev)
    for it in range(4):
        for a, b in zip(pa, pb):
            g = torch.randn_like(a) * (0.1 + it)
            a.grad, b.grad = g.clone(), g.clone()
        oa.step()
        ob.step(gate=gate)
    for a, b in zip(pa, pb):
        assert rel(b, a) < 1e-6
    before = [p.detach().clone() for p in pb]
    ob.step(gate=torch.zeros(1, dtype=torch.int64, device=dev))          # gated off: nothing moves
    assert all(torch.equal(x, y) for x, y in zip(before, pb))
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["param_groups"][0]["params"] == sb["param_groups"][0]["params"]
    for k in sa["state"]:
        assert set(sa["state"][k]) == set(sb["state"][k])
        assert float(sb["state"][k]["step"]) == 4.0
        assert rel(sb["state"][k]["exp_avg_sq"], sa["state"][k]["exp_avg_sq"]) < 1e-6
    oc = torch.optim.AdamW([p.detach().clone().requires_grad_(True) for p in pb], lr=1e-3, weight_decay=1e-4)
    oc.load_state_dict(sb)                                               # checkpoints interchange


def test_checkpoint_round_trip_and_compat(dev):
    from stereo_depth_estimation_b200 import StereoUNet, load_state_dict_compat

    model, sd = fresh_model(dev)
    batch = make_batch(dev, 1, 32, 48)
    model.eval()
    with torch.inference_mode():
        want = model(batch["input"])
    legacy = {k.replace("disparity_head", "output_head"): v.cpu() for k, v in sd.items() if "logvar_head" not in k}
    other = StereoUNet().to(dev)
    missing, unexpected = load_state_dict_compat(other, legacy)
    assert missing == [] and unexpected == []
    other.eval()
    with torch.inference_mode():
        got = other(batch["input"])
    assert torch.equal(got, want)
    with pytest.raises(ValueError):
        model(torch.zeros(1, 6, 30, 48, device=dev))
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 6, 32, 48))


@pytest.mark.gpu
def test_source_prefetcher_delivers_batches_in_order(dev):
    """Row N3: pinned and pageable host batches arrive intact and in order through the double-buffered copy."""
    from stereo_depth_estimation_b200.pipeline import SourcePrefetcher

    g = torch.Generator().manual_seed(5)
    batches = []
    for i in range(5):
        trip = [torch.randint(0, 256, (2, 36, 48, 3), dtype=torch.uint8, generator=g) for _ in range(3)]
        if i % 2 == 0:
            trip = [t.pin_memory() for t in trip]
        batches.append(trip)
    seen = 0
    for i, (left, right, disp, done) in enumerate(SourcePrefetcher(batches, dev)):
        for got, want in zip((left, right, disp), batches[i]):
            assert got.is_cuda and torch.equal(got.cpu(), want)
        done()
        seen += 1
    assert seen == 5


# ------------------------------------------------------------ rows N3 / N4
def test_cached_sample_path_bit_exact(dev):
    """Row N3: the npz-cache sample format (uint8 HWC views + float16 disparity at model resolution,
    dataset.py:86-128) through sdn_preprocess_cached: fixture written by the real reference, then a batch."""
    from stereo_depth_estimation_b200.preprocess import DevicePreprocessor, ViewAug

    f = np.load(os.path.join(GOLDEN, "live_cached.npz"))
    h, w = (int(v) for v in f["cache_hw"])
    pre = DevicePreprocessor(dev, 4, (h, w))
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    args = [torch.from_numpy(f[k][None]).to(dev) for k in ("cache_left", "cache_right", "cache_disp")]
    out = pre.from_cache(*args, count_out=cnt)
    assert np.array_equal(out["input"][0].cpu().numpy(), f["hit_input"])
    assert np.array_equal(out["target"][0].cpu().numpy(), f["hit_target"])
    assert np.array_equal(out["valid_mask"][0].cpu().numpy(), f["hit_mask"])
    assert int(cnt.item()) == int(f["hit_mask"].sum())
    # a ragged batch with augmentation, an all-invalid sample and non-finite cache values
    rng = np.random.default_rng(4)
    b = 3
    L = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
    R = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
    D = (rng.random((b, h, w)) * 70000.0 - 100.0).astype(np.float16)     # negatives, > 65504 -> +inf
    D[1] = 0
    views = [ViewAug(1.1, 0.85, 1.2, 0.05, 0.9), ViewAug(0.8, 1.2, 0.8, -0.09, 1.2, blur_sigma=0.7), ViewAug(), ViewAug(),
             ViewAug(0.9, 1.1, 0.0, -0.5, 1.0), ViewAug(1.0, 1.0, 1.0, 0.0, 1.0, blur_sigma=0.4)]
    dev_args = [torch.from_numpy(a).to(dev) for a in (L, R, D)]
    plain = pre.from_cache(*dev_args, count_out=cnt)
    total = 0
    for i in range(b):
        ref = so.cached_sample(L[i], R[i], D[i])
        assert np.array_equal(plain["input"][i].cpu().numpy(), ref["input"])
        assert np.array_equal(plain["target"][i].cpu().numpy(), ref["target"])
        assert np.array_equal(plain["valid_mask"][i].cpu().numpy(), ref["valid_mask"])
        total += int((ref["valid_mask"] & np.isfinite(ref["target"])).sum())
    assert int(cnt.item()) == total and np.isinf(plain["target"].cpu().numpy()).any()
    aug_out = pre.from_cache(*dev_args, aug=views)
    for i in range(b):
        aug = [dict(brightness=v.brightness, contrast=v.contrast, saturation=v.saturation, hue=v.hue, gamma=v.gamma,
                    blur_sigma=v.blur_sigma) for v in views[2 * i: 2 * i + 2]]
        ref = so.cached_sample(L[i], R[i], D[i], aug=aug)
        np.testing.assert_allclose(aug_out["input"][i].cpu().numpy(), ref["input"], atol=1e-5, rtol=0)
    with pytest.raises(ValueError):
        pre.from_cache(dev_args[0], dev_args[1], dev_args[2].float())


@pytest.mark.parametrize("hs,ws", [(480, 640), (720, 1280), (123, 457), (120, 160)])
def test_live_preprocess_bit_exact(dev, hs, ws):
    """Row N4: preprocess_rgb x 2 + cat (depth_live_dl.py:225-229, 516-520) on the device; OpenCV's uint8
    INTER_LINEAR fixed-point arithmetic must be reproduced bit for bit (oracle, and cv2 itself when present)."""
    from stereo_depth_estimation_b200 import LivePipeline, StereoUNet

    model = StereoUNet().to(dev)
    live = LivePipeline(model, model_size=(320, 240))
    rng = np.random.default_rng(hs)
    view_l = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
    view_r = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
    x = live.preprocess(view_l, view_r).cpu().numpy()
    assert np.array_equal(x, so.live_model_input(view_l, view_r, (320, 240)))
    try:
        import cv2
    except ImportError:
        return
    for v, view in enumerate((view_l, view_r)):
        want = cv2.resize(cv2.cvtColor(view, cv2.COLOR_BGR2RGB), (320, 240), interpolation=cv2.INTER_LINEAR)
        want = (torch.from_numpy(want).float().permute(2, 0, 1) / 255.0).numpy()        # depth_live_dl.py:228
        assert np.array_equal(x[0, 3 * v: 3 * v + 3], want)


def test_live_fixture_and_postprocess(dev):
    """Fixture produced by the real reference (cv2 resize at three scale ratios; disparity_to_depth,
    confidence_from_logvar, depth_live_dl.py:371-381) and the EMA recurrence (:531-538) over three frames."""
    from stereo_depth_estimation_b200 import LivePipeline, StereoUNet

    f = np.load(os.path.join(GOLDEN, "live_cached.npz"))
    model = StereoUNet().to(dev)
    live = LivePipeline(model, model_size=(64, 48), ema_alpha=0.25, focal_length_px=float(f["focal"]),
                        baseline_m=float(f["baseline"]))
    for key in ("f2x", "fragged", "fup"):
        frames = f["frames_" + key]
        x = live.preprocess(frames[0], frames[1]).cpu().numpy()
        assert np.array_equal(x[0, :3], f["pre_" + key][0]) and np.array_equal(x[0, 3:], f["pre_" + key][1]), key
    d0 = f["disparity"]
    lv = torch.from_numpy(f["logvar"]).to(dev)
    smoothed = None
    for frame in range(3):
        d = d0 * np.float32(1.0 + 0.1 * frame)
        maps = live.postprocess(torch.from_numpy(d).to(dev), lv).cpu().numpy()
        smoothed = so.ema_update(smoothed, d, 0.25)
        assert np.array_equal(maps[0], smoothed, equal_nan=True), frame                       # EMA: bit-exact
        assert np.array_equal(maps[1], so.disparity_to_depth(smoothed, float(f["focal"]), float(f["baseline"])),
                              equal_nan=True), frame                                            # one fp32 divide: bit-exact
        np.testing.assert_allclose(maps[2], f["confidence"], rtol=2e-6)                        # expf: ulps
    plain = LivePipeline(model, model_size=(64, 48))            # no EMA, no calibration
    maps = plain.postprocess(torch.from_numpy(d0).to(dev), lv).cpu().numpy()
    assert np.array_equal(maps[0], d0, equal_nan=True) and np.isnan(maps[1]).all()
    live.reset()
    maps = live.postprocess(torch.from_numpy(d0).to(dev), lv).cpu().numpy()
    assert np.array_equal(maps[1], f["depth"], equal_nan=True)


def test_graphed_train_step_matches_eager(dev):
    """GraphedTrainStep (whole step replayed as a CUDA graph) against the same steps launched eagerly: same
    sources, same augmentation parameters, same seed; only the order of the fp32 wgrad atomics differs."""
    from stereo_depth_estimation_b200.optim import FusedAdamW
    from stereo_depth_estimation_b200.preprocess import AugmentSampler, DevicePreprocessor
    from stereo_depth_estimation_b200.step import FusedStep, GraphedTrainStep

    rng = np.random.default_rng(8)
    b, h, w = 2, 64, 96
    srcs = [torch.from_numpy(a).to(dev) for a in synth_sources(rng, b, 108, 192)]
    augs = [AugmentSampler(seed=3).sample_packed(b) for _ in range(7)]
    results = []
    for graphed in (False, False, True):          # two eager runs: their difference is the run-to-run yardstick
        model, sd = fresh_model(dev)
        step = FusedStep(model, FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4))
        pre = DevicePreprocessor(dev, b, (h, w))
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        g = GraphedTrainStep(step, pre) if graphed else None
        out = None
        for aug in augs:
            if g is not None:
                g(srcs[0], srcs[1], srcs[2], aug)
            else:
                out = pre(srcs[0], srcs[1], srcs[2], aug=aug, out=out, count_out=count)
                step.train_step(out, valid_count=count)
        torch.cuda.synchronize()
        if g is not None:
            assert sum(e["graph"] is not None for e in g.entries.values()) == 2      # both parities captured
        results.append(({k: v.detach().clone() for k, v in model.state_dict().items()}, step.read_metrics(), sd))
        # the module still works after replays (operand cache refreshed): eval forward vs the oracle at these weights
        model.eval()
        x = make_batch(dev, 1, h, w, seed=9)["input"]
        with torch.inference_mode():
            d = model(x)
        rd, _ = so.model_forward(results[-1][0], x, False, True)
        assert (d - rd).abs().max().item() / rd.abs().max().item() <= 1e-2
    (pa, ma, sd), (pe, me, _), (pb, mb, _) = results
    assert ma["count"] == mb["count"] == me["count"] > 0
    for k in ("nll", "abs", "sq", "sigma"):      # (sums over seven chaotic training steps: eager-vs-eager differs too)
        assert mb[k] == pytest.approx(ma[k], rel=1e-3), k

    def cosines(p, q):
        out = {}
        for k in p:
            if so.is_param_key(k) and p[k].numel() > 1:
                ua, ub = (p[k] - sd[k]).flatten().double(), (q[k] - sd[k]).flatten().double()
                out[k] = float(torch.dot(ua, ub) / (ua.norm() * ub.norm()).clamp(min=1e-30))
        return out

    # seven AdamW steps amplify the ~1e-7 atomics-order noise of the weight gradients (sign-like first steps):
    # graph-vs-eager must look like eager-vs-eager
    c_graph, c_eager = cosines(pa, pb), cosines(pa, pe)
    worst = min(c_graph, key=c_graph.get)
    mean_graph, mean_eager = sum(c_graph.values()) / len(c_graph), sum(c_eager.values()) / len(c_eager)
    print(f"update cosine graph-vs-eager mean {mean_graph:.4f} min {c_graph[worst]:.4f} ({worst}); "
          f"eager-vs-eager mean {mean_eager:.4f} min {min(c_eager.values()):.4f}")
    # (two EAGER runs of the same seven steps agree to mean 0.985 / worst tensor 0.96-0.99, run to run: the bar for
    # graph-vs-eager is that same neighbourhood)
    # of all parameters taken as ONE vector (dominated by the conv weights, stable run to run) ...
    ua = torch.cat([(pa[k] - sd[k]).flatten().double() for k in c_graph])
    ub = torch.cat([(pb[k] - sd[k]).flatten().double() for k in c_graph])
    ue = torch.cat([(pe[k] - sd[k]).flatten().double() for k in c_graph])
    g_graph = float(torch.dot(ua, ub) / (ua.norm() * ub.norm()))
    g_eager = float(torch.dot(ua, ue) / (ua.norm() * ue.norm()))
    print(f"whole-update cosine graph-vs-eager {g_graph:.4f}, eager-vs-eager {g_eager:.4f}")
    # (measured over many runs: both comparisons land anywhere in 0.983-0.995, so the bar is absolute, not relative
    # to the one eager-vs-eager sample of this run; a replay that used stale operands or skipped a kernel gives < 0.5)
    assert g_graph > 0.95
    # ... and per tensor (the worst one is a 32-512-element BatchNorm bias whose tiny update is mostly noise in
    # BOTH comparisons: 0.96-0.99 eager-vs-eager, so only a gross mismatch is an error there)
    assert mean_graph > 0.95 and c_graph[worst] > 0.8
    for k in pa:
        if "num_batches" in k:
            assert int(pa[k]) == int(pb[k]) == 7
