"""Drop-in ``StereoUNet`` whose forward / backward run on libsdn_b200 (sm_100a).

Boundary replaced (reference: sdfgeoff/stereo_depth_estimation):
  * ``StereoUNet(in_channels=6, out_channels=1, base_channels=32)`` and
    ``forward(x, return_uncertainty=False)``  - src/foundation_stereo_depth/model.py:48-104
  * ``load_state_dict_compat``                 - model.py:8-29
  * callers: train.py:576-578,328,342,431 and live_camera/depth_live_dl.py:394,215,523.

The sub-modules are REAL ``nn.Conv2d / BatchNorm2d / ConvTranspose2d`` objects
registered in the reference's order, so ``state_dict()`` (120 entries),
``parameters()`` (66 tensors, 7,763,938 elements), default initialisation under
``torch.manual_seed`` and AdamW state are identical for free.  Only ``forward``
differs: it hands the parameter storages to the CUDA library, which runs the
whole encoder-decoder with hand-written tcgen05/TMA kernels in bf16 with fp32
accumulation.  There is no PyTorch-op fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_void_p
from typing import Optional

import torch
import torch.nn as nn

from . import _lib


def load_state_dict_compat(model: nn.Module, state_dict: dict) -> tuple[list, list]:
    """Load a checkpoint, accepting the legacy single-head layout (model.py:8-29):
    ``output_head.*`` is renamed to ``disparity_head.*`` and a missing
    ``logvar_head.*`` keeps the model's current values.  Returns
    (missing_keys, unexpected_keys) of a non-strict load."""
    remapped = {}
    for key, value in state_dict.items():
        remapped[key] = value
    for suffix in ("weight", "bias"):
        legacy, current = f"output_head.{suffix}", f"disparity_head.{suffix}"
        if legacy in remapped and current not in remapped:
            remapped[current] = remapped.pop(legacy)
    own = model.state_dict()
    for suffix in ("weight", "bias"):
        key = f"logvar_head.{suffix}"
        if key not in remapped:
            remapped[key] = own[key]
    outcome = model.load_state_dict(remapped, strict=False)
    return list(outcome.missing_keys), list(outcome.unexpected_keys)


class ConvBlock(nn.Module):
    """conv3x3(no bias) -> BN -> ReLU, twice (state_dict keys ``block.{0,1,3,4}.*``,
    reference model.py:32-45).  ``forward`` exists only for completeness of the
    module protocol; StereoUNet never calls it (the CUDA engine runs the block)."""

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        layers = []
        for cin in (in_channels, out_channels):
            layers += [
                nn.Conv2d(cin, out_channels, kernel_size=3, padding=1, bias=False),
                nn.BatchNorm2d(out_channels),
                nn.ReLU(inplace=True),
            ]
        self.block = nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # pragma: no cover - not on the product path
        raise RuntimeError("ConvBlock is executed inside libsdn_b200; call StereoUNet.forward instead")


class _Engine:
    """Owns one sdn_ctx per (device, H, W) and keeps it bound to the module's storages."""

    def __init__(self) -> None:
        self.ctx: Optional[c_void_p] = None
        self.key = None
        self.max_batch = 0
        self.bound_ptrs = None
        self.packed_versions = None
        self.packed_mode = None      # "train" (plain weights + dgrad packing) or "eval" (BatchNorm folded in)
        self.grad_flat: Optional[torch.Tensor] = None
        self.grad_views = None       # gradient destinations currently bound in the context (or None)
        self.graphs = {}             # (batch, want_logvar) -> captured eval forward
        self.generation = 0          # bumped by every forward that overwrites the saved activations

    def close(self) -> None:
        if self.ctx is not None:
            _lib.load().sdn_destroy(self.ctx)
            self.ctx = None

    def __del__(self) -> None:  # best effort
        try:
            self.close()
        except Exception:
            pass

    # a context is a cache, never state: copies / pickles of the module start without one
    def __deepcopy__(self, memo) -> "_Engine":
        return _Engine()

    def __reduce__(self):
        return (_Engine, ())

    def ensure(self, device: torch.device, batch: int, height: int, width: int) -> None:
        lib = _lib.load()
        key = (device.index if device.index is not None else torch.cuda.current_device(), height, width)
        if self.ctx is not None and key == self.key and batch <= self.max_batch:
            return
        self.close()
        ctx = c_void_p()
        _lib.check(lib.sdn_create(ctypes.byref(ctx), key[0], batch, height, width, 0))
        self.ctx, self.key, self.max_batch = ctx, key, batch
        self.bound_ptrs = None
        self.packed_versions = None
        self.packed_mode = None
        self.graphs = {}
        self.grad_views = None


def _ptr_array(tensors) -> ctypes.Array:
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class _StereoFunction(torch.autograd.Function):
    """autograd node for the whole network: one forward, one backward."""

    @staticmethod
    def forward(ctx, module, x, want_logvar, *params):
        disp, logvar = module._launch_forward(x, want_logvar, training=module.training)
        ctx.module = module
        ctx.want_logvar = want_logvar
        # The activations, BatchNorm statistics and pool arg-max of this forward live in the ONE workspace
        # of the module's context: a later forward overwrites them.  Stamp the generation so that a backward
        # through a stale graph raises instead of silently differentiating the newer input.
        eng = module._engine
        ctx.generation = eng.generation      # _launch_forward bumped it
        ctx.ctx_id = eng.ctx.value
        if want_logvar:
            return disp, logvar
        return disp

    @staticmethod
    def backward(ctx, *grad_outputs):
        module = ctx.module
        eng = module._engine
        if eng.ctx is None or ctx.generation != eng.generation or ctx.ctx_id != eng.ctx.value:
            raise RuntimeError(
                "libsdn_b200 keeps the activations of ONE training forward per module: another forward ran "
                "(or the workspace was re-created) between this graph's forward and its backward. Call "
                "backward() before the next forward (gradient accumulation over several backward() calls is fine)."
            )
        g_disp = grad_outputs[0]
        g_logvar = grad_outputs[1] if ctx.want_logvar else None
        if g_disp is None:   # only logvar was used by the loss
            g_disp = torch.zeros_like(g_logvar)
        grads = module._launch_backward(g_disp, g_logvar, ctx.want_logvar)
        return (None, None, None, *grads)


class StereoUNet(nn.Module):
    """5-level U-Net on a 6-channel (left RGB ++ right RGB) input with a softplus
    disparity head and a clamped log-variance head (reference model.py:48-104)."""

    def __init__(self, in_channels: int = 6, out_channels: int = 1, base_channels: int = 32) -> None:
        super().__init__()
        widths = [base_channels * (2**i) for i in range(5)]
        self.pool = nn.MaxPool2d(2)
        # registration order == reference order: enc1..enc4, bottleneck, (up, dec) x4, heads
        self.enc1 = ConvBlock(in_channels, widths[0])
        self.enc2 = ConvBlock(widths[0], widths[1])
        self.enc3 = ConvBlock(widths[1], widths[2])
        self.enc4 = ConvBlock(widths[2], widths[3])
        self.bottleneck = ConvBlock(widths[3], widths[4])
        for level in (4, 3, 2, 1):
            wide, narrow = widths[level], widths[level - 1]
            setattr(self, f"up{level}", nn.ConvTranspose2d(wide, narrow, kernel_size=2, stride=2))
            setattr(self, f"dec{level}", ConvBlock(narrow * 2, narrow))
        self.disparity_head = nn.Conv2d(widths[0], out_channels, kernel_size=1)
        self.logvar_head = nn.Conv2d(widths[0], 1, kernel_size=1)
        self._config = (in_channels, out_channels, base_channels)
        self._engine = _Engine()

    # ------------------------------------------------------------ plumbing
    def _bn_layers(self):
        cached = self.__dict__.get("_bn_cache")
        if cached is None:
            cached = [m for m in self.modules() if isinstance(m, nn.BatchNorm2d)]
            self.__dict__["_bn_cache"] = cached
        return cached

    def _param_list(self):
        """Parameter objects are stable (``.to()`` / ``load_state_dict`` swap ``.data`` in place), so the
        per-call cost of the binding check is 84 ``data_ptr()`` reads instead of a module walk."""
        cached = self.__dict__.get("_param_cache")
        if cached is None:
            cached = list(self.parameters())
            self.__dict__["_param_cache"] = cached
        return cached

    def _check_input(self, x: torch.Tensor) -> None:
        if self._config != (6, 1, 32):
            raise NotImplementedError(
                f"libsdn_b200 implements StereoUNet(6, 1, 32) (the reference's only configuration); got {self._config}"
            )
        if not x.is_cuda:
            raise RuntimeError(
                "B200-native StereoUNet needs CUDA tensors: there is no CPU fallback "
                "(the reference's --device cpu path is the baseline, not the product)"
            )
        if x.dim() != 4 or x.shape[1] != 6:
            raise ValueError(f"expected input [B, 6, H, W], got {tuple(x.shape)}")
        if x.shape[2] % 16 or x.shape[3] % 16:
            raise ValueError(f"H and W must be multiples of 16, got {tuple(x.shape[2:])}")

    def _send_params(self) -> None:
        """Hand the current parameter / BatchNorm-buffer / gradient-destination pointers to the context."""
        lib = _lib.load()
        eng = self._engine
        params = self._param_list()
        bns = self._bn_layers()
        _lib.check(
            lib.sdn_set_params(
                eng.ctx,
                _ptr_array(params),
                _ptr_array(eng.grad_views) if eng.grad_views is not None else None,
                _ptr_array([b.running_mean for b in bns]),
                _ptr_array([b.running_var for b in bns]),
                _ptr_array([b.num_batches_tracked for b in bns]),
            )
        )
        eng.bound_ptrs = self._ptr_key()

    def _ptr_key(self):
        eng = self._engine
        params = tuple(p.data_ptr() for p in self._param_list()) + tuple(b.running_mean.data_ptr() for b in self._bn_layers())
        grads = None if eng.grad_views is None else tuple(None if v is None else v.data_ptr() for v in eng.grad_views)
        return params, grads

    def _bind(self, x: torch.Tensor, training: Optional[bool] = None) -> bool:
        """Make sure the context exists for this shape and points at the current
        parameter storages; returns True when the bf16 operand cache is stale."""
        eng = self._engine
        eng.ensure(x.device, x.shape[0], x.shape[2], x.shape[3])
        for p in self._param_list():
            if p.device != x.device or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("parameters must be contiguous fp32 tensors on the input's device")
        key = self._ptr_key()
        if key != eng.bound_ptrs:
            if eng.bound_ptrs is None or key[0] != eng.bound_ptrs[0]:
                eng.packed_versions = None       # new parameter storages: the operand cache is stale
            self._send_params()
        training = self.training if training is None else training
        dirty = self._versions() != eng.packed_versions or eng.packed_mode != ("train" if training else "eval")
        return dirty

    def _versions(self):
        """Parameter AND BatchNorm-buffer versions: the eval packing folds the running statistics in."""
        bns = self._bn_layers()
        return tuple(p._version for p in self._param_list()) + tuple(b.running_var._version for b in bns) + \
            tuple(b.running_mean._version for b in bns)

    def _pre_forward(self, x: torch.Tensor, training: bool):
        """Common prologue of every library forward: fp32 contiguous input, context bound to the current
        storages.  Returns (x, dirty): dirty = the bf16 operand cache must be re-packed by this call."""
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        dirty = self._bind(x, training)
        self._engine.generation += 1      # this forward overwrites the activations an older graph saved
        return x, dirty

    def _post_forward(self, dirty: bool, training: bool) -> None:
        if dirty:
            eng = self._engine
            eng.packed_versions = self._versions()
            eng.packed_mode = "train" if training else "eval"
            eng.graphs = {}

    def _launch_forward(self, x: torch.Tensor, want_logvar: bool, training: bool, want_outputs: bool = True):
        lib = _lib.load()
        x, dirty = self._pre_forward(x, training)
        eng = self._engine
        b, _, h, w = x.shape
        disp = torch.empty((b, 1, h, w), device=x.device, dtype=torch.float32) if want_outputs else None
        logvar = torch.empty_like(disp) if (want_outputs and want_logvar) else None
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(
            lib.sdn_forward(
                eng.ctx,
                x.data_ptr(),
                disp.data_ptr() if disp is not None else None,
                logvar.data_ptr() if logvar is not None else None,
                b,
                1 if training else 0,
                1 if dirty else 0,
                stream,
            )
        )
        self._post_forward(dirty, training)
        return disp, logvar

    # Small-batch inference (the live viewer's per-frame call, depth_live_dl.py:518-529) is bound by
    # launch latency, not by math: replay the ~28 kernels of the eval forward as ONE CUDA graph.
    GRAPH_MAX_BATCH = 8

    def _forward_eval_graphed(self, x: torch.Tensor, want_logvar: bool):
        eng = self._engine
        x = x.detach().float().contiguous()
        dirty = self._bind(x, False)
        key = (x.shape[0], want_logvar)
        entry = None if dirty else eng.graphs.get(key)
        if entry is None:
            # eager call first: (re)packs the weights if needed and warms the kernels up
            disp, logvar = self._launch_forward(x, want_logvar, training=False)
            static_x = x.clone()
            torch.cuda.current_stream(x.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_disp, g_logvar = self._launch_forward(static_x, want_logvar, training=False)
            eng.graphs[key] = (graph, static_x, g_disp, g_logvar)
            return disp, logvar
        graph, static_x, g_disp, g_logvar = entry
        static_x.copy_(x)
        eng.generation += 1          # the replay overwrites the workspace activations too
        graph.replay()
        return g_disp.clone(), (g_logvar.clone() if g_logvar is not None else None)

    def _new_grad_views(self, device: torch.device):
        params = list(self.parameters())
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, device=device, dtype=torch.float32)
        views, offset = [], 0
        for p in params:
            views.append(flat[offset : offset + p.numel()].view_as(p))
            offset += p.numel()
        return flat, views

    def _bind_grads(self, views) -> None:
        """views: 66 gradient destinations in parameters() order; None entries are not written."""
        self._engine.grad_views = list(views)
        self._send_params()

    def _launch_backward(self, g_disp: torch.Tensor, g_logvar: Optional[torch.Tensor], want_logvar: bool = True):
        lib = _lib.load()
        eng = self._engine
        device = g_disp.device
        flat, views = self._new_grad_views(device)
        if not want_logvar:
            # forward(x) without the uncertainty head: reference autograd leaves logvar_head.{weight,bias}.grad
            # None (so AdamW skips them, no weight decay, no state) - do the same instead of dense zeros
            views[64] = views[65] = None
        self._bind_grads(views)
        g_disp = g_disp.contiguous().float()
        g_logvar = g_logvar.contiguous().float() if g_logvar is not None else None
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(
            lib.sdn_backward_begin(
                eng.ctx, g_disp.data_ptr(), g_logvar.data_ptr() if g_logvar is not None else None, 0, stream
            )
        )
        for stage in range(_lib.NUM_STAGES):
            _lib.check(lib.sdn_backward_stage(eng.ctx, stage, stream))
        eng.grad_flat = flat
        return views

    # ------------------------------------------------------------- forward
    @torch.compiler.disable
    def forward(self, x: torch.Tensor, return_uncertainty: bool = False):
        self._check_input(x)
        params = self._param_list()
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if needs_grad:
            if not self.training:
                raise RuntimeError("gradients through an eval-mode StereoUNet are not implemented in libsdn_b200")
            return _StereoFunction.apply(self, x, bool(return_uncertainty), *params)
        if (not self.training) and x.shape[0] <= self.GRAPH_MAX_BATCH and os.environ.get("SDN_CUDA_GRAPH", "1") != "0":
            disp, logvar = self._forward_eval_graphed(x, bool(return_uncertainty))
        else:
            disp, logvar = self._launch_forward(x, bool(return_uncertainty), training=self.training)
        if return_uncertainty:
            return disp, logvar
        return disp

    # test / debug helper: NHWC activation of conv layer `which` from the last forward
    def debug_activation(self, which: int, kind: int) -> torch.Tensor:
        lib = _lib.load()
        eng = self._engine
        cap = 1 << 28
        dims = (ctypes.c_int * 4)()
        probe = (ctypes.c_float * 1)()
        rc = lib.sdn_debug_read(eng.ctx, which, kind, probe, 0, dims)
        n = dims[0] * dims[1] * dims[2] * dims[3]
        if n <= 0 or n > cap:
            _lib.check(rc)
        buf = torch.empty(n, dtype=torch.float32)
        _lib.check(
            lib.sdn_debug_read(eng.ctx, which, kind, ctypes.cast(buf.data_ptr(), ctypes.POINTER(ctypes.c_float)), n, dims)
        )
        return buf.view(dims[0], dims[1], dims[2], dims[3])

    def profile_enable(self, enable: bool = True) -> None:
        _lib.check(_lib.load().sdn_profile_enable(self._engine.ctx, 1 if enable else 0))

    def profile_dump(self) -> list:
        return _lib.profile_dump(self._engine.ctx)

    def launch_count(self) -> int:
        eng = self._engine
        return int(_lib.load().sdn_launch_count(eng.ctx)) if eng.ctx is not None else 0
