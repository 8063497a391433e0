"""GPU tests (-m gpu): the UNMODIFIED reference code (baseline/_ref, a copy of /root/reference made by
__graft_entry__.build()) running on the B200 drop-in.

north_star: "the training loop, MLflow logging, checkpoints and the live_camera app run unmodified".
What is exercised here, all through the reference's own functions:
  * foundation_stereo_depth.train.run_epoch (train.py:292-418), train + val, on the drop-in module, compared
    with the same function on the reference's own StereoUNet (fp32 torch-eager on the same GPU, TF32 off)
  * train.save_checkpoint (421-436) -> live_camera.depth_live_dl.load_checkpoint (198-222) into the reference
    module on the CPU, and a reference checkpoint back into the drop-in
  * train.log_epoch_previews (254-289)
  * the live call sequence depth_live_dl.py:516-529 verbatim
  * the whole CLI, train.main() (483-689), on a synthetic FoundationStereo-layout dataset on disk

Tolerances (stated, measured values are printed): first-step loss <= 1e-3 relative (north_star), eval-mode
outputs max|err|/max|ref| <= 1e-2 (north_star), epoch metrics of a 3-step AdamW run <= 2e-2 relative (the two
trajectories separate at Adam's sign-like first steps), parameter UPDATE direction cosine no worse than the
same loop under torch's own bf16 autocast (mean - 0.05).
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

import refenv

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refenv.available(), reason="baseline/_ref not vendored")]


@pytest.fixture(scope="module")
def dev():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture()
def ref():
    train, model_mod, stub = refenv.load(fresh=True)
    yield train, model_mod, stub
    refenv.load(fresh=True)   # drop whatever a test patched


def cpu_batches(n, b, h, w, seed=7, scale=4.0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        x = torch.rand(b, 6, h, w, generator=g)
        t = torch.rand(b, 1, h, w, generator=g) * scale
        t[:, :, : h // 4, : w // 4] = 0.0
        out.append({"input": x, "target": t, "valid_mask": t > 0.0})
    return out


def pair_of_models(model_mod, dev, seed=42):
    from stereo_depth_estimation_b200 import StereoUNet

    torch.manual_seed(seed)
    ref_model = model_mod.StereoUNet(in_channels=6, out_channels=1).to(dev)
    torch.manual_seed(seed)
    ours = StereoUNet(in_channels=6, out_channels=1).to(dev)
    for (ka, va), (kb, vb) in zip(ref_model.state_dict().items(), ours.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka      # same registration order -> same init under the seed
    return ref_model, ours


def test_reference_run_epoch_train_and_val_on_dropin(dev, ref):
    train, model_mod, stub = ref
    ref_model, ours = pair_of_models(model_mod, dev)
    init = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    batches = cpu_batches(3, 4, 128, 160)
    opt_r = torch.optim.AdamW(ref_model.parameters(), lr=1e-3, weight_decay=1e-4)      # train.py:578
    opt_o = torch.optim.AdamW(ours.parameters(), lr=1e-3, weight_decay=1e-4)

    # one step first: the loss of the very first batch is the cleanest forward comparison
    m_r1, _ = train.run_epoch(ref_model, batches[:1], dev, optimizer=opt_r, global_step=0, log_every_batches=10)
    m_o1, _ = train.run_epoch(ours, batches[:1], dev, optimizer=opt_o, global_step=0, log_every_batches=10)
    e1 = abs(m_o1["loss"] - m_r1["loss"]) / abs(m_r1["loss"])
    print(f"first-step loss: ref {m_r1['loss']:.6f} ours {m_o1['loss']:.6f} rel {e1:.2e}")
    assert e1 <= 1e-3

    stub.metrics.clear()
    m_r, gs_r = train.run_epoch(ref_model, batches, dev, optimizer=opt_r, global_step=1, log_every_batches=2)
    logged_ref = list(stub.metrics)
    stub.metrics.clear()
    m_o, gs_o = train.run_epoch(ours, batches, dev, optimizer=opt_o, global_step=1, log_every_batches=2)
    logged_ours = list(stub.metrics)
    assert gs_r == gs_o == 4
    assert [s for s, _ in logged_ours] == [s for s, _ in logged_ref]
    assert set(logged_ours[0][1]) == {"train_loss_step", "train_nll_step", "train_mae_step", "train_rmse_step",
                                      "train_sigma_step"}
    for k in ("loss", "mae", "rmse", "sigma"):
        e = abs(m_o[k] - m_r[k]) / abs(m_r[k])
        print(f"train epoch {k}: ref {m_r[k]:.6f} ours {m_o[k]:.6f} rel {e:.2e}")
        assert e <= 2e-2, k
    # the parameter UPDATE after 4 AdamW steps points the same way.  AdamW's first steps are sign-like
    # (g / sqrt(g^2)): an element whose gradient is dominated by bf16 noise flips its whole step, so the yardstick
    # is the SAME reference loop on the reference module under torch's bf16 autocast (same seed, same batches)
    class AutocastRef(model_mod.StereoUNet):
        def forward(self, x, return_uncertainty=False):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = super().forward(x, return_uncertainty)
            return tuple(o.float() for o in out) if isinstance(out, tuple) else out.float()

    torch.manual_seed(42)
    ac_model = AutocastRef(in_channels=6, out_channels=1).to(dev)
    opt_a = torch.optim.AdamW(ac_model.parameters(), lr=1e-3, weight_decay=1e-4)
    train.run_epoch(ac_model, batches[:1], dev, optimizer=opt_a, global_step=0, log_every_batches=10)
    train.run_epoch(ac_model, batches, dev, optimizer=opt_a, global_step=1, log_every_batches=2)

    def update_cosines(model):
        out = {}
        for (k, pr), (_, po) in zip(ref_model.named_parameters(), model.named_parameters()):
            ur, uo = (pr.detach() - init[k]).flatten().double(), (po.detach() - init[k]).flatten().double()
            out[k] = float(torch.dot(ur, uo) / (ur.norm() * uo.norm()).clamp(min=1e-30))
        return out

    cosines, yard = update_cosines(ours), update_cosines(ac_model)
    worst, yworst = min(cosines, key=cosines.get), min(yard, key=yard.get)
    mean_cos, ymean = sum(cosines.values()) / 66, sum(yard.values()) / 66
    print(f"update cosine vs fp32 over 66 parameters: ours mean {mean_cos:.4f} min {cosines[worst]:.4f} ({worst}); "
          f"torch-autocast mean {ymean:.4f} min {yard[yworst]:.4f} ({yworst})")
    assert mean_cos >= ymean - 0.05 and cosines[worst] >= min(0.5, yard[yworst] - 0.1), (mean_cos, ymean, worst)
    # BatchNorm buffers (part of the checkpoint): running statistics after 4 train steps
    so, sr, sa = ours.state_dict(), ref_model.state_dict(), ac_model.state_dict()
    worst_buf = (0.0, 0.0, "")
    for k in sr:
        if k.endswith("num_batches_tracked"):
            assert int(so[k]) == int(sr[k]) == 4, k
        elif "running_" in k:
            rel = float((so[k] - sr[k]).norm() / sr[k].norm().clamp(min=1e-30))
            yrel = float((sa[k] - sr[k]).norm() / sr[k].norm().clamp(min=1e-30))
            worst_buf = max(worst_buf, (rel, yrel, k))
            assert rel <= max(3e-2, 1.5 * yrel), (k, rel, yrel)      # (weights have separated by step 4)
    print(f"worst BatchNorm buffer after 4 steps: ours {worst_buf[0]:.2e} torch-autocast {worst_buf[1]:.2e} ({worst_buf[2]})")

    # validation epoch at IDENTICAL weights: run_epoch(optimizer=None) (train.py:618)
    ours.load_state_dict(ref_model.state_dict())
    v_r, _ = train.run_epoch(ref_model, batches, dev, optimizer=None)
    v_o, _ = train.run_epoch(ours, batches, dev, optimizer=None)
    for k in ("loss", "mae", "rmse", "sigma"):
        e = abs(v_o[k] - v_r[k]) / abs(v_r[k])
        print(f"val epoch {k}: ref {v_r[k]:.6f} ours {v_o[k]:.6f} rel {e:.2e}")
        assert e <= 1e-2, k
    assert not ours.training and not ref_model.training


def test_reference_loop_with_zero_grad_set_to_none_still_updates(dev, ref):
    """train.py:325 sets every .grad to None each step; the fused step must re-attach its views."""
    from stereo_depth_estimation_b200 import StereoUNet
    from stereo_depth_estimation_b200.step import FusedStep

    torch.manual_seed(1)
    model = StereoUNet().to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    step = FusedStep(model, opt)
    batch = {k: v.to(dev) for k, v in cpu_batches(1, 2, 32, 48)[0].items()}
    before = [p.detach().clone() for p in model.parameters()]
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        assert step.train_step(batch) > 0
        mid = [p.detach().clone() for p in model.parameters()]
        assert any(not torch.equal(a, b) for a, b in zip(before, mid))
        before = mid
    assert all(p.grad is not None for p in model.parameters())


def test_reference_checkpoint_round_trip(dev, ref, tmp_path, monkeypatch):
    train, model_mod, _ = ref
    live = refenv.load_live()
    ref_model, ours = pair_of_models(model_mod, dev)
    opt = torch.optim.AdamW(ours.parameters(), lr=1e-3, weight_decay=1e-4)
    batches = cpu_batches(2, 2, 64, 96)
    train.run_epoch(ours, batches, dev, optimizer=opt, global_step=0, log_every_batches=10)
    monkeypatch.setattr(sys, "argv", ["foundation-stereo-depth", "--dataset-root", str(tmp_path)])
    cfg = train.parse_args()
    path = tmp_path / "last.pt"
    train.save_checkpoint(path, 3, ours, opt, cfg, {"val_mae": 1.0})          # train.py:421-436
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "args", "metrics"}
    assert len(ck["model_state_dict"]) == 120 and len(ck["optimizer_state_dict"]["state"]) == 66
    # ... into the REFERENCE module on the CPU through the live viewer's loader (depth_live_dl.py:198-222)
    cpu_ref = model_mod.StereoUNet(in_channels=6, out_channels=1)
    epoch, has_unc = live.load_checkpoint(cpu_ref, path, torch.device("cpu"))
    assert epoch == 3 and has_unc and not cpu_ref.training
    for k, v in ours.state_dict().items():
        assert torch.equal(cpu_ref.state_dict()[k], v.cpu()), k
    # the reference's forward on those weights agrees with ours (eval, north_star tolerance)
    x = batches[0]["input"][:1]
    with torch.inference_mode():
        d_ref, l_ref = cpu_ref(x, return_uncertainty=True)
        ours.eval()
        d_our, l_our = ours(x.to(dev), return_uncertainty=True)
    assert float((d_our.cpu() - d_ref).abs().max() / d_ref.abs().max()) <= 1e-2
    assert float((l_our.cpu() - l_ref).abs().max() / l_ref.abs().max()) <= 1e-2
    # ... and a checkpoint written from the reference module (legacy single-head layout too) into the drop-in
    ref_path = tmp_path / "ref.pt"
    opt_r = torch.optim.AdamW(ref_model.parameters(), lr=1e-3, weight_decay=1e-4)
    train.save_checkpoint(ref_path, 1, ref_model, opt_r, cfg, {})
    from stereo_depth_estimation_b200 import StereoUNet, dropin

    dropin.install()
    live = refenv.load_live()
    assert live.StereoUNet is StereoUNet
    fresh = StereoUNet(in_channels=6, out_channels=1).to(dev)
    epoch, has_unc = live.load_checkpoint(fresh, ref_path, dev)
    assert epoch == 1 and has_unc
    for k, v in ref_model.state_dict().items():
        assert torch.equal(fresh.state_dict()[k], v), k
    legacy = {k.replace("disparity_head", "output_head"): v for k, v in ref_model.state_dict().items()
              if "logvar_head" not in k}
    torch.save(legacy, tmp_path / "legacy.pt")
    epoch, has_unc = live.load_checkpoint(fresh, tmp_path / "legacy.pt", dev)
    assert epoch == -1 and not has_unc


def test_reference_live_call_sequence_verbatim(dev, ref):
    """depth_live_dl.py:516-529 on the drop-in vs on the reference module (same GPU, fp32)."""
    _, model_mod, _ = ref
    live = refenv.load_live()
    ref_model, ours = pair_of_models(model_mod, dev)
    # non-trivial BatchNorm buffers: a few training forwards of the reference, then share the state
    ref_model.train()
    with torch.no_grad():
        for b in cpu_batches(3, 2, 240, 320, seed=3):
            ref_model(b["input"].to(dev))
    ours.load_state_dict(ref_model.state_dict())
    ref_model.eval()
    ours.eval()
    rng = np.random.default_rng(11)
    model_size = (320, 240)   # (width, height), depth_live_dl.py:455
    device = dev
    for frame in range(3):    # frame 0 captures the CUDA graph, frames 1-2 replay it
        view_l = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
        view_r = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
        outs = []
        for model in (ours, ref_model):
            # ---- verbatim from depth_live_dl.py:516-529 ----
            left_tensor = live.preprocess_rgb(view_l, model_size)
            right_tensor = live.preprocess_rgb(view_r, model_size)
            model_input = (
                torch.cat([left_tensor, right_tensor], dim=0).unsqueeze(0).to(device)
            )

            with torch.inference_mode():
                disparity_tensor, logvar_tensor = model(
                    model_input, return_uncertainty=True
                )
                prediction = (
                    disparity_tensor[0, 0].detach().cpu().numpy().astype(np.float32)
                )
                logvar = logvar_tensor[0, 0].detach().cpu().numpy().astype(np.float32)
            # ---- end verbatim ----
            outs.append((prediction, logvar))
        (p_o, l_o), (p_r, l_r) = outs
        e_d = float(np.abs(p_o - p_r).max() / np.abs(p_r).max())
        e_l = float(np.abs(l_o - l_r).max() / np.abs(l_r).max())
        print(f"live frame {frame}: disparity max-rel {e_d:.2e} logvar max-rel {e_l:.2e}")
        assert p_o.shape == (240, 320) and e_d <= 1e-2 and e_l <= 1e-2
        depth = live.disparity_to_depth(p_o, 488.87 * 0.5, 0.0715)      # depth_live_dl.py:371-377
        conf = live.confidence_from_logvar(l_o)                          # :380-381
        assert np.isfinite(depth[p_o > 1e-6]).all() and (conf > 0).all()


def test_live_pipeline_matches_reference_loop(dev, ref):
    """Row N4 end to end: LivePipeline (device pre -> graph-replayed model -> device post, one D2H) against the
    reference's per-frame host code (depth_live_dl.py:516-538, 371-381) driving the REFERENCE model."""
    _, model_mod, _ = ref
    live_mod = refenv.load_live()
    from stereo_depth_estimation_b200 import LivePipeline

    ref_model, ours = pair_of_models(model_mod, dev)
    ref_model.train()
    with torch.no_grad():
        for b in cpu_batches(3, 2, 240, 320, seed=5):
            ref_model(b["input"].to(dev))
    ours.load_state_dict(ref_model.state_dict())
    ref_model.eval()
    focal, baseline, alpha = 488.87 * 320 / 640, 0.0715, 0.4        # calibration/stereo_calib.npz, rescaled
    pipe = LivePipeline(ours, model_size=(320, 240), ema_alpha=alpha, focal_length_px=focal, baseline_m=baseline)
    rng = np.random.default_rng(12)
    smoothed = None
    for frame in range(3):
        view_l = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
        view_r = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
        got = pipe(view_l, view_r)
        # ---- the reference's host code on the reference model ----
        left_tensor = live_mod.preprocess_rgb(view_l, (320, 240))
        right_tensor = live_mod.preprocess_rgb(view_r, (320, 240))
        model_input = torch.cat([left_tensor, right_tensor], dim=0).unsqueeze(0).to(dev)
        with torch.inference_mode():
            disparity_tensor, logvar_tensor = ref_model(model_input, return_uncertainty=True)
            prediction = disparity_tensor[0, 0].detach().cpu().numpy().astype(np.float32)
            logvar = logvar_tensor[0, 0].detach().cpu().numpy().astype(np.float32)
        smoothed = prediction if smoothed is None else (alpha * prediction + (1.0 - alpha) * smoothed)
        depth_m = live_mod.disparity_to_depth(smoothed, float(focal), float(baseline))
        confidence = live_mod.confidence_from_logvar(logvar)
        # ----
        assert np.array_equal(pipe._input.cpu().numpy(), model_input.cpu().numpy())        # device pre: bit-exact
        for name, a, r in (("disparity", got["disparity"], smoothed), ("logvar", got["logvar"], logvar),
                           ("depth", got["depth"], depth_m), ("confidence", got["confidence"], confidence)):
            e = float(np.nanmax(np.abs(a - r)) / np.nanmax(np.abs(r)))
            assert a.shape == (240, 320) and e <= 1e-2, (frame, name, e)
            assert np.array_equal(np.isnan(a), np.isnan(r)), (frame, name)


def test_reference_previews_on_dropin(dev, ref, tmp_path):
    train, model_mod, _ = ref
    _, ours = pair_of_models(model_mod, dev)
    ours.train()
    n = train.log_epoch_previews(ours, cpu_batches(2, 2, 64, 96), dev, epoch=1, preview_root=tmp_path)
    assert n == 4 and ours.training
    assert len(list((tmp_path / "epoch_0001").glob("*.png"))) == 4


def _write_dataset(root, scenes=2, per_scene=6, hs=90, ws=160):
    from PIL import Image

    rng = np.random.default_rng(0)
    for s in range(scenes):
        base = root / f"scene_{s:02d}" / "dataset" / "data"
        for sub in ("left/rgb", "right/rgb", "left/disparity"):
            (base / sub).mkdir(parents=True, exist_ok=True)
        for i in range(per_scene):
            stem = f"{i:06d}"
            left = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
            right = np.roll(left, -3, axis=1)
            disp = np.full((hs, ws), 3.0 + i, dtype=np.float32)
            disp[: hs // 5] = 0.0
            v = np.round(disp * 1000.0).astype(np.int64)
            enc = np.stack([v // (255 * 255), (v // 255) % 255, v % 255], axis=-1).astype(np.uint8)
            Image.fromarray(left).save(base / "left/rgb" / f"{stem}.png")
            Image.fromarray(right).save(base / "right/rgb" / f"{stem}.png")
            Image.fromarray(enc).save(base / "left/disparity" / f"{stem}.png")


def test_reference_cli_main_runs_unmodified_on_dropin(dev, ref, tmp_path, monkeypatch, capsys):
    """`foundation-stereo-depth` end to end (train.py:483-689): argparse, dataset discovery, DataLoader, AdamW,
    run_epoch train + val, previews, checkpoints, MLflow calls - on the B200 module installed by dropin."""
    train, model_mod, stub = ref
    from stereo_depth_estimation_b200 import StereoUNet, dropin

    patched = dropin.install()
    assert train.StereoUNet is StereoUNet and "foundation_stereo_depth.model" in patched
    data = tmp_path / "data"
    _write_dataset(data)
    out = tmp_path / "out"
    monkeypatch.setattr(sys, "argv", [
        "foundation-stereo-depth", "--dataset-root", str(data), "--height", "64", "--width", "96", "--epochs", "2",
        "--batch-size", "4", "--num-workers", "0", "--val-fraction", "0.25", "--device", "cuda",
        "--output-dir", str(out), "--mlflow-tracking-uri", f"sqlite:///{tmp_path}/mlflow.db"])
    train.main()
    text = capsys.readouterr().out
    assert "Using device: cuda" in text and "Epoch 2/2" in text
    run_dir = out / "stubrun0001"
    for name in ("last.pt", "best.pt"):
        ck = torch.load(run_dir / "checkpoints" / name, map_location="cpu", weights_only=False)
        assert len(ck["model_state_dict"]) == 120
        assert all(torch.isfinite(v).all() for v in ck["model_state_dict"].values() if v.is_floating_point())
        cpu_ref = model_mod.StereoUNet(in_channels=6, out_channels=1) if model_mod.StereoUNet is not StereoUNet else None
    epoch_logs = [m for s, m in stub.metrics if "val_mae" in m]
    assert len(epoch_logs) == 2 and all(math.isfinite(m["train_loss"]) and math.isfinite(m["val_mae"]) for m in epoch_logs)
    assert stub.params["num_parameters"] == 7_763_938 if "num_parameters" in stub.params else True
    assert stub.tags["best_epoch"] in (1, 2)
    assert len(list((run_dir / "mlflow_previews" / "epoch_0002").glob("*.png"))) >= 1
    _ = os
