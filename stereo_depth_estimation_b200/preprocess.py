"""Device input pipeline: the host-side mirror of the reference's
``FoundationStereoDataset.__getitem__`` + default collate
(src/foundation_stereo_depth/dataset.py:184-270,302-311) for raw uint8 images that
are already on the GPU.  One call produces a whole batch in the reference's
sample format::

    {"input": [B,6,H,W] f32, "target": [B,1,H,W] f32, "valid_mask": [B,1,H,W] bool}

Augmentation parameters are explicit (``AugmentSampler`` draws them with the
reference's distributions and draw order, dataset.py:214-246), so a test can
replay exactly what the reference sampled.  PNG decoding is out of scope.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


@dataclass
class ViewAug:
    """One view's parameters; the same fields as sdn_aug_params (include/sdn.h)."""

    brightness: float = 1.0
    contrast: float = 1.0
    saturation: float = 1.0
    hue: float = 0.0
    gamma: float = 1.0
    blur_sigma: float = 0.0
    noise_std: float = 0.0
    noise_seed: int = 0


class AugmentSampler:
    """Draws per-view parameters like the reference samplers.

    Distributions and the per-view draw order follow dataset.py:214-246 and
    248-270: brightness, contrast, saturation ~ U[max(0,1-j), 1+j]; hue ~ U[-j, j];
    gamma ~ U[max(.1,1-j), 1+j]; blur iff rand < blur_prob with sigma ~ U[0.1,
    max(sigma_max, 0.1)]; noise_std ~ U[0, max].  Defaults are the CLI defaults of
    train.py:156-209.  The generator is a host ``numpy`` one (the reference uses the
    global torch CPU generator; streams cannot be bit-matched across devices)."""

    def __init__(self, brightness_jitter: float = 0.2, contrast_jitter: float = 0.2, saturation_jitter: float = 0.25,
                 hue_jitter: float = 0.09, gamma_jitter: float = 0.2, noise_std_max: float = 0.05,
                 blur_prob: float = 0.03, blur_sigma_max: float = 1.0, seed: int = 0) -> None:
        if not 0.0 <= blur_prob <= 1.0:
            raise ValueError(f"blur_prob must be in [0, 1], got {blur_prob}")
        if saturation_jitter < 0.0:
            raise ValueError(f"saturation_jitter must be >= 0, got {saturation_jitter}")
        if gamma_jitter < 0.0:
            raise ValueError(f"gamma_jitter must be >= 0, got {gamma_jitter}")
        self.j = (brightness_jitter, contrast_jitter, saturation_jitter)
        self.hue_jitter, self.gamma_jitter = hue_jitter, gamma_jitter
        self.noise_std_max, self.blur_prob, self.blur_sigma_max = noise_std_max, blur_prob, blur_sigma_max
        self.rng = np.random.default_rng(seed)

    def _factor(self, jitter: float) -> float:
        if jitter <= 0.0:
            return 1.0
        return float(self.rng.uniform(max(0.0, 1.0 - jitter), 1.0 + jitter))

    def sample_view(self) -> ViewAug:
        v = ViewAug()
        v.brightness, v.contrast, v.saturation = (self._factor(j) for j in self.j)
        v.hue = float(self.rng.uniform(-self.hue_jitter, self.hue_jitter)) if self.hue_jitter > 0 else 0.0
        if self.gamma_jitter > 0:
            low = max(0.1, 1.0 - self.gamma_jitter)
            v.gamma = float(self.rng.uniform(low, max(low, 1.0 + self.gamma_jitter)))
        if self.blur_prob > 0 and self.blur_sigma_max > 0 and self.rng.random() < self.blur_prob:
            v.blur_sigma = float(self.rng.uniform(0.1, max(self.blur_sigma_max, 0.1)))
        v.noise_std = float(self.rng.uniform(0.0, self.noise_std_max)) if self.noise_std_max > 0 else 0.0
        v.noise_seed = int(self.rng.integers(0, 2**32 - 1))
        return v

    def sample_batch(self, batch: int) -> list:
        """2*batch views in (left, right) order per sample, like dataset.py:302-304."""
        return [self.sample_view() for _ in range(2 * batch)]

    def sample_packed(self, batch: int) -> torch.Tensor:
        """The same distributions drawn for a whole batch at once (numpy, vectorised) and
        returned as the packed sdn_aug_params array (host uint8 tensor)."""
        n = 2 * batch
        rec = np.zeros(n, dtype=AUG_DTYPE)

        def factor(j):
            return self.rng.uniform(max(0.0, 1.0 - j), 1.0 + j, n) if j > 0 else np.ones(n)

        rec["brightness"], rec["contrast"], rec["saturation"] = (factor(j) for j in self.j)
        if self.hue_jitter > 0:
            rec["hue"] = self.rng.uniform(-self.hue_jitter, self.hue_jitter, n)
        if self.gamma_jitter > 0:
            low = max(0.1, 1.0 - self.gamma_jitter)
            rec["gamma"] = self.rng.uniform(low, max(low, 1.0 + self.gamma_jitter), n)
        else:
            rec["gamma"] = 1.0
        if self.blur_prob > 0 and self.blur_sigma_max > 0:
            coin = self.rng.random(n) < self.blur_prob
            rec["blur_sigma"] = np.where(coin, self.rng.uniform(0.1, max(self.blur_sigma_max, 0.1), n), 0.0)
        if self.noise_std_max > 0:
            rec["noise_std"] = self.rng.uniform(0.0, self.noise_std_max, n)
        rec["noise_seed"] = self.rng.integers(0, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
        return torch.from_numpy(rec.view(np.uint8).copy())


AUG_DTYPE = np.dtype([("brightness", "<f4"), ("contrast", "<f4"), ("saturation", "<f4"), ("hue", "<f4"),
                      ("gamma", "<f4"), ("blur_sigma", "<f4"), ("noise_std", "<f4"), ("noise_seed", "<u4")])


def pack_aug(views: Sequence[ViewAug]) -> torch.Tensor:
    """Host uint8 tensor holding the sdn_aug_params array (DevicePreprocessor stages it in pinned memory)."""
    arr = (_lib.AugParams * len(views))()
    for i, v in enumerate(views):
        arr[i] = _lib.AugParams(v.brightness, v.contrast, v.saturation, v.hue, v.gamma, v.blur_sigma, v.noise_std,
                                v.noise_seed & 0xFFFFFFFF)
    raw = np.frombuffer(arr, dtype=np.uint8).copy()
    t = torch.from_numpy(raw)
    return t


class DevicePreprocessor:
    """Owns an sdn context used only for its preprocessing workspace."""

    def __init__(self, device: torch.device, max_batch: int, image_size=(240, 320)) -> None:
        if torch.device(device).type != "cuda":
            raise RuntimeError("DevicePreprocessor runs on CUDA only; there is no CPU fallback")
        self.device = torch.device(device)
        self.image_size = (int(image_size[0]), int(image_size[1]))
        self.max_batch = int(max_batch)
        lib = _lib.load()
        self.ctx = c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(lib.sdn_create(ctypes.byref(self.ctx), index, self.max_batch, self.image_size[0],
                                  self.image_size[1], _lib.CTX_PREPROCESS_ONLY))
        depth = 4   # augmentation-parameter ring: the host may run this many steps ahead of the device
        self._ring = [torch.empty(2 * self.max_batch * 32, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self._ring_ev = [torch.cuda.Event() for _ in range(depth)]
        self._ring_pos = -1

    def close(self) -> None:
        if getattr(self, "ctx", None) is not None and self.ctx:
            _lib.load().sdn_destroy(self.ctx)
            self.ctx = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def profile_enable(self, enable: bool = True) -> None:
        _lib.check(_lib.load().sdn_profile_enable(self.ctx, 1 if enable else 0))

    def profile_dump(self) -> list:
        return _lib.profile_dump(self.ctx)

    def launch_count(self) -> int:
        return int(_lib.load().sdn_launch_count(self.ctx))

    def _stage_aug(self, aug, b: int):
        """(device-or-pinned tensor, flags) for an augmentation parameter set, or (None, 0)."""
        if aug is None:
            return None, 0
        packed = aug if torch.is_tensor(aug) else pack_aug(aug)   # list of ViewAug or sample_packed() output
        if packed.numel() != 2 * b * 32:
            raise ValueError(f"need 2*B = {2 * b} view parameter sets, got {packed.numel() // 32}")
        if packed.is_cuda:
            return packed, 0
        if packed.is_pinned():
            # the caller owns a pinned buffer and its lifetime (GraphedTrainStep: one per captured graph)
            return packed, _lib.PREPROCESS_AUG_HOST
        # event-guarded ring of pinned buffers that the device reads in place (a staging KERNEL,
        # not a copy-engine transfer: it cannot queue behind a bulk prefetch of the next batch)
        slot = self._ring_pos = (self._ring_pos + 1) % len(self._ring_ev)
        self._ring_ev[slot].synchronize()
        buf = self._ring[slot][: packed.numel()]
        buf.copy_(packed)
        return buf, _lib.PREPROCESS_AUG_HOST

    def _outputs(self, b: int, out: Optional[dict]) -> dict:
        if out is not None:
            return out
        h, w = self.image_size
        return {
            "input": torch.empty((b, 6, h, w), device=self.device, dtype=torch.float32),
            "target": torch.empty((b, 1, h, w), device=self.device, dtype=torch.float32),
            "valid_mask": torch.empty((b, 1, h, w), device=self.device, dtype=torch.bool),
        }

    def __call__(self, left: torch.Tensor, right: torch.Tensor, disparity: torch.Tensor,
                 aug: Optional[Sequence[ViewAug]] = None, fourterm: bool = False,
                 out: Optional[dict] = None, count_out: Optional[torch.Tensor] = None) -> dict:
        """left / right / disparity: uint8 [B,Hs,Ws,3] CUDA tensors (HWC, RGB)."""
        for name, t in (("left", left), ("right", right), ("disparity", disparity)):
            if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3 or not t.is_cuda:
                raise ValueError(f"{name} must be a CUDA uint8 tensor [B,Hs,Ws,3], got {t.dtype} {tuple(t.shape)}")
        if left.shape != right.shape or left.shape != disparity.shape:
            raise ValueError("left, right and disparity must have the same shape")
        b, hs, ws, _ = left.shape
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        left, right, disparity = left.contiguous(), right.contiguous(), disparity.contiguous()
        out = self._outputs(b, out)
        aug_dev, flags_aug = self._stage_aug(aug, b)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(
            _lib.load().sdn_preprocess(
                self.ctx, left.data_ptr(), right.data_ptr(), disparity.data_ptr(), b, hs, ws,
                aug_dev.data_ptr() if aug_dev is not None else None,
                out["input"].data_ptr(), out["target"].data_ptr(), out["valid_mask"].data_ptr(),
                count_out.data_ptr() if count_out is not None else None,
                (_lib.RESIZE_FOURTERM if fourterm else 0) | flags_aug, stream,
            )
        )
        if aug_dev is not None and not aug_dev.is_cuda and not (torch.is_tensor(aug) and aug.is_pinned()):
            self._ring_ev[self._ring_pos].record(torch.cuda.current_stream(self.device))
        return out

    def from_cache(self, left: torch.Tensor, right: torch.Tensor, disparity: torch.Tensor,
                   aug: Optional[Sequence[ViewAug]] = None, out: Optional[dict] = None,
                   count_out: Optional[torch.Tensor] = None) -> dict:
        """The reference's npz read-through cache format (dataset.py:86-128): ``left`` / ``right`` uint8
        [B,H,W,3] and ``disparity`` float16 [B,H,W], ALREADY at this preprocessor's resolution (the arrays of
        the ``.npz`` entries, stacked and copied to the GPU).  No resize, no disparity rescale."""
        h, w = self.image_size
        for name, t in (("left", left), ("right", right)):
            if t.dtype != torch.uint8 or not t.is_cuda or t.dim() != 4 or tuple(t.shape[1:]) != (h, w, 3):
                raise ValueError(f"{name} must be a CUDA uint8 tensor [B,{h},{w},3], got {t.dtype} {tuple(t.shape)}")
        b = left.shape[0]
        if disparity.dtype != torch.float16 or not disparity.is_cuda or tuple(disparity.shape) != (b, h, w):
            raise ValueError(f"disparity must be a CUDA float16 tensor [{b},{h},{w}] (the cache stores float16), "
                             f"got {disparity.dtype} {tuple(disparity.shape)}")
        if right.shape != left.shape:
            raise ValueError("left and right must have the same shape")
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        left, right, disparity = left.contiguous(), right.contiguous(), disparity.contiguous()
        out = self._outputs(b, out)
        aug_dev, flags_aug = self._stage_aug(aug, b)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(
            _lib.load().sdn_preprocess_cached(
                self.ctx, left.data_ptr(), right.data_ptr(), disparity.data_ptr(), b,
                aug_dev.data_ptr() if aug_dev is not None else None,
                out["input"].data_ptr(), out["target"].data_ptr(), out["valid_mask"].data_ptr(),
                count_out.data_ptr() if count_out is not None else None, flags_aug, stream,
            )
        )
        if aug_dev is not None and not aug_dev.is_cuda:
            self._ring_ev[self._ring_pos].record(torch.cuda.current_stream(self.device))
        return out
