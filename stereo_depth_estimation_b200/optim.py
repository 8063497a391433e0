"""FusedAdamW: ``torch.optim.AdamW`` semantics (the reference's optimizer, train.py:578: lr 1e-3,
betas (0.9, 0.999), eps 1e-8, decoupled weight decay 1e-4) as ONE kernel launch over all
parameter tensors (``sdn_adamw_step``), with the reference's "no valid pixel -> no optimizer
step" rule (train.py:331-332) evaluated on the device, so the train step needs no host sync.

``state_dict()`` keeps torch's layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so a
checkpoint written by train.py:429-436 with this optimizer loads into ``torch.optim.AdamW`` and back.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from typing import Optional

import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step_dev = {}   # group index -> device int64 step counter
        self._ctx: Optional[c_void_p] = None

    def _context(self, device: torch.device) -> c_void_p:
        if self._ctx is None:
            ctx = c_void_p()
            index = device.index if device.index is not None else torch.cuda.current_device()
            _lib.check(_lib.load().sdn_create(ctypes.byref(ctx), index, 1, 16, 16, _lib.CTX_PREPROCESS_ONLY))
            self._ctx = ctx
        return self._ctx

    def __del__(self):
        try:
            if self._ctx is not None:
                _lib.load().sdn_destroy(self._ctx)
        except Exception:
            pass

    @torch.no_grad()
    def step(self, closure=None, gate: Optional[torch.Tensor] = None):
        """``gate``: optional device int64/uint64 tensor [1]; the update is skipped on the device when it is 0."""
        loss = closure() if closure is not None else None
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            device = params[0].device
            if device.type != "cuda":
                raise RuntimeError("FusedAdamW runs on CUDA tensors only")
            for start in range(0, len(params), 66):
                chunk = params[start:start + 66]
                for p in chunk:
                    st = self.state[p]
                    if not st:
                        st["step"] = torch.tensor(0.0)
                        st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                        st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                        raise RuntimeError("FusedAdamW needs contiguous fp32 parameters and gradients")
                key = (gi, start)
                if key not in self._step_dev:
                    first = self.state[chunk[0]]["step"]
                    self._step_dev[key] = torch.full((1,), int(float(first)), dtype=torch.int64, device=device)
                n = len(chunk)
                arr = lambda ts: (c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
                numel = (ctypes.c_int64 * n)(*[p.numel() for p in chunk])
                b1, b2 = group["betas"]
                _lib.check(lib.sdn_adamw_step(
                    self._context(device), arr(chunk), arr([p.grad for p in chunk]),
                    arr([self.state[p]["exp_avg"] for p in chunk]), arr([self.state[p]["exp_avg_sq"] for p in chunk]),
                    numel, n, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                    float(group["weight_decay"]), self._step_dev[key].data_ptr(),
                    gate.data_ptr() if gate is not None else None,
                    torch.cuda.current_stream(device).cuda_stream))
                # the library wrote the parameters behind torch's back: bump their version counters so
                # that StereoUNet re-packs its bf16 operand cache (and autograd sees the mutation)
                for p in chunk:
                    torch.autograd.graph.increment_version(p)
        return loss

    def _sync_steps(self) -> None:
        for (gi, start), counter in self._step_dev.items():
            value = float(counter.item())
            params = [p for p in self.param_groups[gi]["params"] if p in self.state]
            for p in params[start:start + 66]:
                self.state[p]["step"] = torch.tensor(value)

    def state_dict(self):
        self._sync_steps()   # the applied-step count lives on the device between checkpoints
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._step_dev = {}
