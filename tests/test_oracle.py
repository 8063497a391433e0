"""Pin the oracle (oracle/stereo_oracle.py) against the reference: the golden
vectors of the reference's own tests (tests/test_dataset.py:17-61) and fixtures
produced by the real reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import stereo_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def g(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


# ---- reference KATs (tests/test_dataset.py:31-35, :38-61) -------------------
def test_decode_round_trip_kat():
    disparity = np.array([[0.0, 0.125, 1.25], [2.0, 3.5, 10.0]], dtype=np.float32)
    decoded = so.decode_disparity(so.encode_disparity(disparity))
    np.testing.assert_allclose(decoded, disparity, atol=1e-3)


def test_decode_max_value():
    out = so.decode_disparity(np.array([[[255, 255, 255]]], dtype=np.uint8))
    assert out.dtype == np.float32
    np.testing.assert_allclose(out, [[16646.654]], rtol=1e-7)


def test_disparity_resize_scales_with_output_width_kat():
    src = np.full((2, 4), 1.5, dtype=np.float32)
    target = so.load_disparity(so.encode_disparity(src), (2, 8))
    assert target.shape == (1, 2, 8)
    np.testing.assert_allclose(target, np.full((1, 2, 8), 3.0, np.float32), atol=1e-3)


# ---- fixtures from the real reference ---------------------------------------
@pytest.mark.parametrize("name", ["sample_exact", "sample_ragged", "sample_up", "sample_same"])
def test_sample_bit_exact(name):
    f = g(name + ".npz")
    out_hw = tuple(int(v) for v in f["out_hw"])
    np.testing.assert_array_equal(so.decode_disparity(f["disp"]), f["decoded"])
    # small images go through torch's "fourterm" instantiation of the CPU kernel: bit-exact
    item = so.make_sample(f["left"], f["right"], f["disp"], out_hw, formula="fourterm")
    np.testing.assert_array_equal(item["input"], f["input"])
    np.testing.assert_array_equal(item["target"], f["target"])
    np.testing.assert_array_equal(item["valid_mask"], f["valid_mask"])
    # the canonical "separable" form differs from it by at most one rounding
    sep = so.make_sample(f["left"], f["right"], f["disp"], out_hw)
    np.testing.assert_allclose(sep["input"], f["input"], atol=1.2e-7, rtol=0)
    np.testing.assert_allclose(sep["target"], f["target"], rtol=2.4e-7, atol=0)


def test_bilinear_matches_torch_interpolate_on_cpu():
    """torch (the dependency the reference calls at dataset.py:187-192) run live on CPU:
    each case must be bit-equal to one of the two instantiations, and the single-channel
    headline case (the disparity plane, 540x960 -> 240x320) to the canonical one."""
    rng = np.random.default_rng(3)
    for (c, hs, ws, h, w) in [(3, 54, 96, 24, 32), (3, 37, 53, 16, 32), (3, 135, 240, 60, 80), (3, 20, 30, 48, 64),
                              (1, 540, 960, 240, 320), (3, 540, 960, 240, 320), (1, 500, 700, 160, 480)]:
        img = rng.random((c, hs, ws), dtype=np.float32)
        ref = torch.nn.functional.interpolate(torch.from_numpy(img)[None], size=(h, w), mode="bilinear",
                                              align_corners=False)[0].numpy()
        sep = so.bilinear_resize(img, (h, w), "separable")
        four = so.bilinear_resize(img, (h, w), "fourterm")
        assert np.array_equal(sep, ref) or np.array_equal(four, ref), (c, hs, ws, h, w)
        np.testing.assert_allclose(sep, four, atol=1.2e-7, rtol=0)
        if c == 1 and hs >= 500:
            np.testing.assert_array_equal(sep, ref)


def test_augment_matches_reference():
    f = g("augment.npz")
    for params, want in zip(f["params"], f["outs"]):
        b, c, s, h, gm, sig = (float(v) for v in params)
        got = so.augment_rgb(f["img"], b, c, s, h, gm, blur_sigma=sig)
        np.testing.assert_allclose(got, want, atol=2e-6, rtol=0)


def _batch(f):
    return {
        "input": torch.from_numpy(f["input"]),
        "target": torch.from_numpy(f["target"]),
        "valid_mask": torch.from_numpy(f["valid_mask"]),
    }


def test_model_forward_loss_grads_match_reference():
    f = g("model.npz")
    sd = so.init_state_dict(42)
    batch = _batch(f)
    names = [str(n) for n in f["param_names"]]
    assert names == so.param_keys(sd)
    sd_eval = {k: v.clone() for k, v in sd.items()}
    sums, grads = so.train_step(sd, batch, so.AdamWState(lr=0.0, weight_decay=0.0))
    disp, logvar = so.model_forward(so.init_state_dict(42), batch["input"], True, True, {})
    np.testing.assert_allclose(disp.numpy(), f["disp_train"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(logvar.numpy(), f["logvar_train"], atol=1e-5, rtol=1e-5)
    assert sums["nll"] / sums["count"] == pytest.approx(float(f["loss"]), rel=1e-5)
    for name, want in zip(names, f["grad_digests"]):
        gr = grads[name].double().flatten()
        idx = torch.linspace(0, gr.numel() - 1, steps=8).long()
        got = np.concatenate([[gr.norm().item(), gr.sum().item()], gr[idx].numpy()])
        np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-6, err_msg=name)
    # BatchNorm buffers after one training forward
    for k in sd:
        if "running" in k or "num_batches" in k:
            np.testing.assert_allclose(sd[k].numpy(), f["bn__" + k.replace(".", "__")], rtol=1e-5, atol=1e-6)
    # eval mode with the UPDATED buffers (the golden eval pass ran after the training pass)
    for k in sd:
        if "running" in k or "num_batches" in k:
            sd_eval[k] = sd[k]
    disp_e, logvar_e = so.model_forward(sd_eval, batch["input"], False, True)
    np.testing.assert_allclose(disp_e.numpy(), f["disp_eval"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(logvar_e.numpy(), f["logvar_eval"], atol=1e-5, rtol=1e-5)


def test_loss_closed_form_matches_autograd():
    torch.manual_seed(0)
    zd = (torch.randn(2, 1, 8, 12) * 3).requires_grad_(True)
    zl = (torch.randn(2, 1, 8, 12) * 4).requires_grad_(True)
    tgt = torch.rand(2, 1, 8, 12) * 5
    tgt[0, 0, 0, 0] = float("nan")
    mask = torch.rand(2, 1, 8, 12) > 0.2
    disp = torch.nn.functional.softplus(zd)
    lv = zl.clamp(-6.0, 3.0)
    loss, _ = so.loss_and_sums(disp, lv, tgt, mask)
    loss.backward()
    gd, gl = so.loss_grads_closed_form(zd.detach(), zl.detach(), tgt, mask)
    np.testing.assert_allclose(gd.numpy(), zd.grad.numpy(), atol=1e-7)
    np.testing.assert_allclose(gl.numpy(), zl.grad.numpy(), atol=1e-7)


def test_empty_mask_is_skipped():
    sd = so.init_state_dict(1)
    batch = {"input": torch.rand(1, 6, 32, 32), "target": torch.zeros(1, 1, 32, 32),
             "valid_mask": torch.zeros(1, 1, 32, 32, dtype=torch.bool)}
    before = sd["enc1.block.0.weight"].clone()
    sums, grads = so.train_step(sd, batch, so.AdamWState())
    assert sums["count"] == 0 and grads == {}
    assert torch.equal(before, sd["enc1.block.0.weight"])
    with pytest.raises(RuntimeError):
        so.run_epoch(sd, [batch], so.AdamWState())


def test_run_epoch_matches_reference():
    f = g("epoch.npz")
    sd = so.init_state_dict(42)
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import synth_batch

    batches = [synth_batch(200 + i, 2, 32, 48) for i in range(3)]
    opt = so.AdamWState()
    train = so.run_epoch(sd, batches, opt)
    val = so.run_epoch(sd, batches[:1], None)
    keys = ("loss", "nll", "mae", "rmse", "sigma")
    np.testing.assert_allclose([train[k] for k in keys], f["train_metrics"], rtol=2e-4)
    np.testing.assert_allclose([val[k] for k in keys], f["val_metrics"], rtol=2e-3)
    assert opt.step == int(f["global_step"])
    for key, want in zip(so.param_keys(sd), f["param_digests"]):
        p = sd[key].double().flatten()
        assert p.norm().item() == pytest.approx(want[0], rel=1e-3), key


# ------------------------------------------------------------------ rows N3 / N4
def test_cached_sample_matches_reference_cache_hit():
    """dataset.py:86-128, 272-311: the fixture holds the arrays the REAL reference wrote into its npz cache and
    the sample it returned on the cache hit."""
    f = g("live_cached.npz")
    got = so.cached_sample(f["cache_left"], f["cache_right"], f["cache_disp"])
    assert f["cache_disp"].dtype == np.float16 and f["cache_left"].dtype == np.uint8
    assert np.array_equal(got["input"], f["hit_input"])
    assert np.array_equal(got["target"], f["hit_target"])
    assert np.array_equal(got["valid_mask"], f["hit_mask"])
    # and the writer side: what save_cached_sample stores for the freshly computed sample
    l8, r8, d16 = so.to_cache_format(f["miss_input"][:3], f["miss_input"][3:], f["hit_target"])
    assert np.array_equal(l8, f["cache_left"]) and np.array_equal(r8, f["cache_right"])
    assert np.array_equal(d16, f["cache_disp"])


def test_live_pre_post_match_reference_fixture():
    """live_camera/depth_live_dl.py:225-229, 371-381 through the fixture produced by the real reference (cv2)."""
    f = g("live_cached.npz")
    for key in ("f2x", "fragged", "fup"):
        frames = f["frames_" + key]
        for v in range(2):
            assert np.array_equal(so.live_preprocess_rgb(frames[v], (64, 48)), f["pre_" + key][v]), (key, v)
        x = so.live_model_input(frames[0], frames[1], (64, 48))
        assert x.shape == (1, 6, 48, 64) and x.dtype == np.float32
    depth = so.disparity_to_depth(f["disparity"], float(f["focal"]), float(f["baseline"]))
    assert np.array_equal(depth, f["depth"], equal_nan=True)
    assert np.isnan(depth[:5]).all() and np.isnan(depth[30, 7]) and np.isnan(depth[31, 7])
    np.testing.assert_allclose(so.confidence_from_logvar(f["logvar"]), f["confidence"], rtol=1e-6)
    p = f["disparity"]
    assert so.ema_update(None, p, 0.4) is p and so.ema_update(p, p, 0.0) is p
    np.testing.assert_array_equal(so.ema_update(p * 2, p, 0.25), (0.25 * p + 0.75 * (p * 2)).astype(np.float32))


def test_cv_resize_restatement_is_bit_exact_against_cv2():
    """The third-party algorithm itself: cv2.resize(INTER_LINEAR) on uint8 (opencv-python 4.13), many shapes."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    shapes = [((480, 640), (320, 240)), ((720, 1280), (320, 240)), ((240, 320), (640, 480)), ((123, 457), (320, 240)),
              ((481, 641), (320, 240)), ((100, 100), (333, 77)), ((7, 9), (64, 48)), ((33, 47), (31, 29))]
    for _ in range(8):
        shapes.append(((int(rng.integers(5, 200)), int(rng.integers(5, 200))),
                       (int(rng.integers(4, 160)), int(rng.integers(4, 160)))))
    for (h, w), (dw, dh) in shapes:
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(so.cv_resize_linear_u8(src, (dw, dh)), want), ((h, w), (dw, dh))
