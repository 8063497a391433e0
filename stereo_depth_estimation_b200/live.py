"""Device-side mirror of the live viewer's per-frame path.

The reference's loop (src/live_camera/depth_live_dl.py:516-538, 371-381) does, per frame and on the host:
``preprocess_rgb`` twice (BGR -> RGB, ``cv2.resize`` INTER_LINEAR on uint8, float / 255, CHW), ``cat``, a 1.84 MB
float32 host -> device copy, the model, two device -> host copies, optional EMA smoothing, ``disparity_to_depth``
and ``confidence_from_logvar``.  ``LivePipeline`` keeps everything between the raw camera frames and the final
maps on the GPU: the two uint8 frames go up (pinned staging), ``sdn_live_preprocess`` builds the model input with
OpenCV's fixed-point resize arithmetic bit-exactly, the model runs (CUDA-graph replay), ``sdn_live_postprocess``
applies EMA / depth / confidence in one pass, and ONE pinned copy brings the maps back::

    live = LivePipeline(model, model_size=(320, 240), ema_alpha=0.4, focal_length_px=244.4, baseline_m=0.0715)
    maps = live(view_l, view_r)      # BGR uint8 numpy frames -> {"disparity", "logvar", "depth", "confidence"}

Colormaps, contours and the HUD stay on the host with OpenCV (GUI code, out of scope)."""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .model import StereoUNet


class LivePipeline:
    def __init__(self, model: StereoUNet, model_size: Tuple[int, int] = (320, 240), device: Optional[torch.device] = None,
                 ema_alpha: float = 0.0, focal_length_px: Optional[float] = None, baseline_m: Optional[float] = None) -> None:
        """``model_size`` = (width, height) like depth_live_dl.py:455; ``focal_length_px`` is the focal length at the
        MODEL resolution (the reference rescales the calibration's by model_width / calibration_width)."""
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("LivePipeline runs on CUDA only; the reference's --device cpu path is the baseline")
        self.width, self.height = int(model_size[0]), int(model_size[1])
        if self.width % 16 or self.height % 16:
            raise ValueError(f"model size must be multiples of 16, got {model_size}")
        self.ema_alpha = float(ema_alpha)
        self.focal_length_px, self.baseline_m = focal_length_px, baseline_m
        self.ctx = c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.load().sdn_create(ctypes.byref(self.ctx), index, 1, self.height, self.width,
                                          _lib.CTX_PREPROCESS_ONLY))
        self._frames_host = None            # pinned uint8 [2, Hs, Ws, 3]
        self._frames_dev = None
        n = self.height * self.width
        self._input = torch.empty((1, 6, self.height, self.width), device=self.device, dtype=torch.float32)
        self._ema = torch.empty(n, device=self.device, dtype=torch.float32)
        self._ema_valid = False
        # planes: (smoothed) disparity, depth, confidence; planes the kernel never writes (no calibration / no
        # uncertainty head) stay NaN from here
        self._maps = torch.full((3, self.height, self.width), float("nan"), device=self.device, dtype=torch.float32)
        self._maps_host = torch.empty((4, self.height, self.width), dtype=torch.float32).pin_memory()

    def close(self) -> None:
        if getattr(self, "ctx", None) is not None and self.ctx:
            _lib.load().sdn_destroy(self.ctx)
            self.ctx = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        """Forget the EMA state (the reference restarts smoothing when the loop restarts)."""
        self._ema_valid = False

    def preprocess(self, view_l, view_r) -> torch.Tensor:
        """Two BGR uint8 frames [Hs,Ws,3] (numpy or tensors, host or device) -> model input [1,6,H,W] float32."""
        lib = _lib.load()
        frames = []
        for v in (view_l, view_r):
            t = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
            if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[-1] != 3:
                raise ValueError(f"frames must be uint8 [Hs,Ws,3] (BGR), got {t.dtype} {tuple(t.shape)}")
            frames.append(t)
        if frames[0].shape != frames[1].shape:
            raise ValueError("left and right frames must have the same shape")
        hs, ws = int(frames[0].shape[0]), int(frames[0].shape[1])
        if all(f.is_cuda for f in frames):
            dev_l, dev_r = frames[0].contiguous(), frames[1].contiguous()
        else:
            if self._frames_host is None or tuple(self._frames_host.shape[1:3]) != (hs, ws):
                self._frames_host = torch.empty((2, hs, ws, 3), dtype=torch.uint8).pin_memory()
                self._frames_dev = torch.empty((2, hs, ws, 3), dtype=torch.uint8, device=self.device)
            self._frames_host[0].copy_(frames[0])
            self._frames_host[1].copy_(frames[1])
            self._frames_dev.copy_(self._frames_host, non_blocking=True)
            dev_l, dev_r = self._frames_dev[0], self._frames_dev[1]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.sdn_live_preprocess(self.ctx, dev_l.data_ptr(), dev_r.data_ptr(), hs, ws, self._input.data_ptr(),
                                           stream))
        return self._input

    def postprocess(self, disparity: torch.Tensor, logvar: Optional[torch.Tensor]) -> torch.Tensor:
        """-> device tensor [3,H,W]: (smoothed) disparity, depth (NaN where invalid, or everywhere when no
        calibration was given), confidence (NaN without a logvar)."""
        lib = _lib.load()
        n = self.height * self.width
        depth_on = self.focal_length_px is not None and self.baseline_m is not None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        maps = self._maps
        _lib.check(lib.sdn_live_postprocess(
            self.ctx, disparity.contiguous().data_ptr(), logvar.contiguous().data_ptr() if logvar is not None else None, n,
            self._ema.data_ptr(), 1 if self._ema_valid else 0, self.ema_alpha,
            float(self.focal_length_px) if depth_on else 0.0, float(self.baseline_m) if depth_on else 0.0,
            maps[0].data_ptr(), maps[1].data_ptr() if depth_on else None,
            maps[2].data_ptr() if logvar is not None else None, stream))
        if self.ema_alpha > 0.0:
            self._ema_valid = True
        return maps

    @torch.inference_mode()
    def __call__(self, view_l, view_r) -> Dict[str, np.ndarray]:
        x = self.preprocess(view_l, view_r)
        self.model.eval()
        disparity, logvar = self.model(x, return_uncertainty=True)
        maps = self.postprocess(disparity, logvar)
        self._maps_host[:3].copy_(maps, non_blocking=True)
        self._maps_host[3].copy_(logvar.reshape(self.height, self.width), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        out = self._maps_host.numpy()
        return {"disparity": out[0].copy(), "depth": out[1].copy(), "confidence": out[2].copy(), "logvar": out[3].copy()}
