"""Generate the golden fixtures in tests/golden/ by running the REAL reference.

Run in the build container only (it imports /root/reference/src, which does not
exist on the GPU box):   python tests/golden/make_golden.py
The fixtures it writes are committed; tests never import the reference.

What is captured (reference = sdfgeoff/stereo_depth_estimation):
  sample_*.npz   FoundationStereoDataset.__getitem__ on PNG triplets written to a
                 temp dir (dataset.py:272-311) -> input / target / valid_mask
  augment.npz    FoundationStereoDataset._augment_rgb (dataset.py:248-270) with the
                 samplers (dataset.py:214-246) pinned to fixed values
  model.npz      StereoUNet (model.py) seed 42: train- and eval-mode outputs, loss,
                 per-parameter gradient digests, updated BatchNorm buffers
  epoch.npz      run_epoch (train.py:292-418) + AdamW over two synthetic batches
  live_cached.npz  rows N3 / N4: save_cached_sample -> load_cached_sample -> __getitem__ on a cache hit
                 (dataset.py:86-128, 272-311), and the live viewer's preprocess_rgb (cv2 fixed-point resize),
                 disparity_to_depth, confidence_from_logvar (live_camera/depth_live_dl.py:225-229, 371-381)
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
from PIL import Image

REF_SRC = "/root/reference/src"
OUT = Path(__file__).resolve().parent


def import_reference():
    sys.path.insert(0, REF_SRC)
    stub = types.ModuleType("mlflow")
    stub.log_metrics = lambda *a, **k: None
    sys.modules.setdefault("mlflow", stub)
    from foundation_stereo_depth import dataset as ref_dataset  # noqa: E402
    from foundation_stereo_depth import model as ref_model  # noqa: E402
    from foundation_stereo_depth import train as ref_train  # noqa: E402

    return ref_dataset, ref_model, ref_train


def synth_triplet(rng: np.random.Generator, hs: int, ws: int):
    left = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
    right = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
    disp = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
    disp[..., 0] = rng.integers(0, 4, (hs, ws), dtype=np.uint8)  # keep disparity <= ~260 px
    disp[rng.random((hs, ws)) < 0.1] = 0  # invalid pixels
    return left, right, disp


def make_samples(ref_dataset) -> None:
    rng = np.random.default_rng(7)
    cases = {
        "sample_exact": (54, 96, 24, 32),   # same ratios as 540x960 -> 240x320 (2.25, 3.0)
        "sample_ragged": (50, 70, 16, 48),  # inexact scales 3.125, 1.4583
        "sample_up": (12, 16, 32, 48),      # upsampling
        "sample_same": (16, 32, 16, 32),    # identity size
    }
    for name, (hs, ws, h, w) in cases.items():
        left, right, disp = synth_triplet(rng, hs, ws)
        with tempfile.TemporaryDirectory() as tmp:
            tmp = Path(tmp)
            Image.fromarray(left, mode="RGB").save(tmp / "l.png")
            Image.fromarray(right, mode="RGB").save(tmp / "r.png")
            Image.fromarray(disp, mode="RGB").save(tmp / "d.png")
            sample = ref_dataset.StereoSample(tmp / "l.png", tmp / "r.png", tmp / "d.png")
            ds = ref_dataset.FoundationStereoDataset([sample], image_size=(h, w))
            item = ds[0]
        np.savez_compressed(
            OUT / f"{name}.npz",
            left=left, right=right, disp=disp, out_hw=np.array([h, w]),
            input=item["input"].numpy(), target=item["target"].numpy(), valid_mask=item["valid_mask"].numpy(),
            decoded=ref_dataset.depth_uint8_decoding(disp),
        )


def make_augment(ref_dataset) -> None:
    rng = np.random.default_rng(11)
    img = rng.random((3, 24, 32), dtype=np.float32)
    img[:, :4, :4] = 0.5  # a gray patch: maxc == minc branch of rgb->hsv
    img[:, 4:6, :4] = 0.0
    cases = [
        dict(brightness=1.1, contrast=0.85, saturation=1.2, hue=0.05, gamma=0.9, blur_sigma=0.0),
        dict(brightness=0.8, contrast=1.2, saturation=0.8, hue=-0.09, gamma=1.2, blur_sigma=0.7),
        dict(brightness=1.0, contrast=1.0, saturation=1.0, hue=0.0, gamma=1.0, blur_sigma=0.0),
        dict(brightness=1.2, contrast=0.8, saturation=1.25, hue=0.5, gamma=0.8, blur_sigma=1.0),
    ]
    sample = ref_dataset.StereoSample(Path("l"), Path("r"), Path("d"))
    outs = []
    for case in cases:
        ds = ref_dataset.FoundationStereoDataset(
            [sample], augment=True, brightness_jitter=0.2, contrast_jitter=0.2, saturation_jitter=0.25,
            hue_jitter=0.09, gamma_jitter=0.2, noise_std_max=0.0, blur_prob=1.0, blur_sigma_max=1.0,
        )
        seq = iter([case["brightness"], case["contrast"], case["saturation"]])
        ds._sample_jitter_factor = lambda jitter, seq=seq: next(seq)
        ds._sample_hue_shift = lambda case=case: case["hue"]
        ds._sample_gamma_factor = lambda case=case: case["gamma"]
        ds._should_apply_blur = lambda case=case: case["blur_sigma"] > 0.0
        ds._sample_blur_sigma = lambda case=case: case["blur_sigma"]
        ds._sample_noise_std = lambda: 0.0
        outs.append(ds._augment_rgb(torch.from_numpy(img.copy())).numpy())
    np.savez_compressed(
        OUT / "augment.npz", img=img, outs=np.stack(outs),
        params=np.array([[c["brightness"], c["contrast"], c["saturation"], c["hue"], c["gamma"], c["blur_sigma"]] for c in cases], dtype=np.float64),
    )


def synth_batch(seed: int, b: int, h: int, w: int):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 6, h, w, generator=g)
    t = torch.rand(b, 1, h, w, generator=g) * 64.0
    t[:, :, : h // 4, : w // 4] = 0.0  # an invalid block
    return {"input": x, "target": t, "valid_mask": t > 0.0}


def digest(t: torch.Tensor) -> np.ndarray:
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, steps=8).long()
    return np.concatenate([[f.norm().item(), f.sum().item()], f[idx].numpy()])


def make_model(ref_model) -> None:
    torch.manual_seed(42)
    model = ref_model.StereoUNet(in_channels=6, out_channels=1)
    batch = synth_batch(123, 2, 32, 48)
    model.train()
    disp, logvar = model(batch["input"], return_uncertainty=True)
    mask = batch["valid_mask"] & torch.isfinite(batch["target"])
    diff = disp[mask] - batch["target"][mask]
    nll = diff.abs() * torch.exp(-logvar[mask]) + logvar[mask]
    loss = nll.mean()
    loss.backward()
    names = [n for n, _ in model.named_parameters()]
    grad_digests = np.stack([digest(p.grad) for _, p in model.named_parameters()])
    bn_after = {k.replace(".", "__"): v.numpy() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
    model.eval()
    with torch.inference_mode():
        disp_e, logvar_e = model(batch["input"], return_uncertainty=True)
        disp_only = model(batch["input"])
    np.savez_compressed(
        OUT / "model.npz",
        input=batch["input"].numpy(), target=batch["target"].numpy(), valid_mask=batch["valid_mask"].numpy(),
        disp_train=disp.detach().numpy(), logvar_train=logvar.detach().numpy(), loss=np.float64(loss.item()),
        grad_digests=grad_digests, param_names=np.array(names),
        disp_eval=disp_e.numpy(), logvar_eval=logvar_e.numpy(), disp_only_equal=np.array(torch.equal(disp_only, disp_e)),
        **{"bn__" + k: v for k, v in bn_after.items()},
    )


def make_epoch(ref_model, ref_train) -> None:
    torch.manual_seed(42)
    model = ref_model.StereoUNet(in_channels=6, out_channels=1)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    batches = [synth_batch(200 + i, 2, 32, 48) for i in range(3)]
    metrics, step = ref_train.run_epoch(model, batches, torch.device("cpu"), optimizer=opt, global_step=0, log_every_batches=10)
    val, _ = ref_train.run_epoch(model, batches[:1], torch.device("cpu"), optimizer=None)
    param_digests = np.stack([digest(p) for p in model.parameters()])
    np.savez_compressed(
        OUT / "epoch.npz",
        train_metrics=np.array([metrics[k] for k in ("loss", "nll", "mae", "rmse", "sigma")]),
        val_metrics=np.array([val[k] for k in ("loss", "nll", "mae", "rmse", "sigma")]),
        global_step=np.array(step), param_digests=param_digests,
    )


def make_live_cached(ref_dataset) -> None:
    from live_camera import depth_live_dl as live  # the reference's viewer module (needs cv2)

    rng = np.random.default_rng(21)
    # N3: a sample written to and read back from the reference's npz cache
    left, right, disp = synth_triplet(rng, 54, 96)
    h, w = 32, 48
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        base = tmp / "scene_00" / "dataset" / "data"
        for sub in ("left/rgb", "right/rgb", "left/disparity"):
            (base / sub).mkdir(parents=True)
        Image.fromarray(left, mode="RGB").save(base / "left/rgb/000001.png")
        Image.fromarray(right, mode="RGB").save(base / "right/rgb/000001.png")
        Image.fromarray(disp, mode="RGB").save(base / "left/disparity/000001.png")
        samples = ref_dataset.discover_samples(tmp)
        cache = tmp / "cache"
        ds = ref_dataset.FoundationStereoDataset(samples, image_size=(h, w), cache_root=cache)
        first = ds[0]                                   # miss: computes and writes the cache entry
        entry = cache / ref_dataset.sample_cache_relpath(samples[0])
        with np.load(entry) as z:
            c_left, c_right, c_disp = z["left"], z["right"], z["disparity"]
        hit = ref_dataset.FoundationStereoDataset(samples, image_size=(h, w), cache_root=cache, require_cache=True)[0]
    # N4: live pre / post on synthetic camera frames
    # (small frames keep the fixture small; the arithmetic does not depend on the size: 2.0x down, a ragged
    # ratio, and an upsample; model size (64, 48) as (width, height))
    frames = {"f2x": rng.integers(0, 256, (2, 96, 128, 3), dtype=np.uint8),
              "fragged": rng.integers(0, 256, (2, 61, 77, 3), dtype=np.uint8),
              "fup": rng.integers(0, 256, (2, 24, 32, 3), dtype=np.uint8)}
    pre = {k: np.stack([live.preprocess_rgb(v[0], (64, 48)).numpy(), live.preprocess_rgb(v[1], (64, 48)).numpy()])
           for k, v in frames.items()}
    d = (rng.random((48, 64)).astype(np.float32) * 40.0)
    d[:5] = 0.0
    d[30, 7] = np.nan
    d[31, 7] = 5e-7
    lv = rng.random((48, 64)).astype(np.float32) * 9.0 - 6.0
    np.savez_compressed(
        OUT / "live_cached.npz",
        cache_left=c_left, cache_right=c_right, cache_disp=c_disp, cache_hw=np.array([h, w]),
        miss_input=first["input"].numpy(), hit_input=hit["input"].numpy(), hit_target=hit["target"].numpy(),
        hit_mask=hit["valid_mask"].numpy(),
        **{"frames_" + k: v for k, v in frames.items()}, **{"pre_" + k: v for k, v in pre.items()},
        disparity=d, logvar=lv, focal=np.float64(244.435), baseline=np.float64(0.0715),
        depth=live.disparity_to_depth(d, 244.435, 0.0715), confidence=live.confidence_from_logvar(lv),
    )


def main() -> None:
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    ref_dataset, ref_model, ref_train = import_reference()
    if "--only-live-cached" in sys.argv:
        make_live_cached(ref_dataset)
        return
    make_samples(ref_dataset)
    make_augment(ref_dataset)
    make_model(ref_model)
    make_epoch(ref_model, ref_train)
    make_live_cached(ref_dataset)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
