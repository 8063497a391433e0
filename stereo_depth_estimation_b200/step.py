"""Fused train / validation step and a ``run_epoch`` with the reference's signature.

The reference computes the loss with ~25 torch ops and 12 ``.item()`` syncs per
step inside ``run_epoch`` (src/foundation_stereo_depth/train.py:320-363).  An
unmodified ``run_epoch`` therefore cannot reach the fused head+loss kernel; this
module is the second entry point (level B2):

  * ``FusedStep.train_step(batch)``  : forward -> fused heteroscedastic loss + head
    backward -> staged backward -> (bucketed all-reduce) -> optimizer step, with
    ONE host sync (the valid-pixel count, needed for the reference's
    "skip the batch when no pixel is valid" rule, train.py:331-332).
  * ``run_epoch(model, loader, device, optimizer, global_step, log_every_batches)``
    : same arguments, same return value and the same MLflow metric names as
    train.py:292-418, driven by ``FusedStep``.

Data parallelism (one process per GPU, ``torch.distributed``): the batch is
sharded by the caller; the loss normaliser is the GLOBAL valid count (all-reduced
while the forward runs) so the sum of per-rank gradients equals the
single-process gradient exactly; the 4 backward stages are the all-reduce
buckets, issued on a side stream as soon as each stage's kernels are enqueued so
NCCL overlaps the rest of the backward.  BatchNorm statistics stay per rank
(standard DDP semantics).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import _lib
from .model import StereoUNet
from .optim import FusedAdamW

try:  # optional: only needed when world_size > 1
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None

MLFLOW_TRAIN_LOG_EVERY_BATCHES = 10  # train.py:23


def stage_slices(model: StereoUNet) -> list:
    """(start, end) element ranges of the flat gradient buffer per backward stage."""
    sizes = [p.numel() for p in model.parameters()]
    offsets = [0]
    for s in sizes:
        offsets.append(offsets[-1] + s)
    out = []
    for stage in range(_lib.NUM_STAGES):
        first, num = _lib.stage_param_range(stage)
        out.append((offsets[first], offsets[first + num]))
    return out


class FusedStep:
    def __init__(self, model: StereoUNet, optimizer: Optional[torch.optim.Optimizer] = None,
                 process_group=None, overlap: bool = True) -> None:
        self.model = model
        self.optimizer = optimizer
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist is not None and dist.is_initialized()) else 1
        self.overlap = overlap
        self.flat = None
        self.views = None
        self.slices = None
        self.sums = None       # device fp32 [4]: sum nll, |diff|, diff^2, exp(.5 logvar)
        self.count = None      # device i64 [1]: valid pixels accumulated with the sums
        self.n_norm = None     # device i64 [1]: loss normaliser of the current step
        self.comm_stream = None

    # -------------------------------------------------------------- buffers
    def _ensure(self, device: torch.device) -> None:
        if self.flat is not None and self.flat.device == device:
            return
        self.flat, self.views = self.model._new_grad_views(device)
        self.flat.zero_()
        self.slices = stage_slices(self.model)
        self.sums = torch.zeros(4, device=device, dtype=torch.float32)
        self.count = torch.zeros(1, device=device, dtype=torch.int64)
        self.n_norm = torch.zeros(1, device=device, dtype=torch.int64)
        if self.world > 1:
            self.comm_stream = torch.cuda.Stream(device=device)
        for p, v in zip(self.model.parameters(), self.views):
            p.grad = v

    def reset_metrics(self) -> None:
        if self.sums is not None:
            self.sums.zero_()
            self.count.zero_()

    def read_metrics(self) -> Dict[str, float]:
        """One D2H copy of the five running sums (train.py:345-357)."""
        s = self.sums.double().cpu()
        n = int(self.count.cpu().item())
        return {"nll": float(s[0]), "abs": float(s[1]), "sq": float(s[2]), "sigma": float(s[3]), "count": n}

    # ----------------------------------------------------------------- steps
    def train_step(self, batch: Dict[str, torch.Tensor], valid_count: Optional[torch.Tensor] = None) -> int:
        """One optimisation step on an already-assembled batch (the reference's sample
        format).  ``valid_count`` (device int64 [1]) may come from the preprocessing
        kernel; otherwise it is counted here.  Returns the (global) valid count; 0 means
        the step was skipped like train.py:331-332 (-1 with ``FusedAdamW``: the rule is applied on the
        device and the host never learns the count)."""
        model = self.model
        lib = _lib.load()
        x = batch["input"]
        target = batch["target"].contiguous()
        mask = batch["valid_mask"].contiguous()
        device = x.device
        self._ensure(device)
        model.train(True)
        main = torch.cuda.current_stream(device)
        stream = main.cuda_stream

        # forward (no head store: the loss kernel recomputes the 1x1 heads from dec1)
        model._check_input(x)
        model._launch_forward(x, True, training=True, want_outputs=False)
        eng = model._engine
        model._bind_grads(self.views)

        if valid_count is None:
            _lib.check(lib.sdn_count_valid(eng.ctx, target.data_ptr(), mask.data_ptr(), x.shape[0],
                                           self.n_norm.data_ptr(), stream))
        else:
            torch.add(valid_count.view(1), 0, out=self.n_norm)   # a kernel, not a copy-engine transfer
        if self.world > 1:
            dist.all_reduce(self.n_norm, group=self.group)

        _lib.check(lib.sdn_loss_begin(eng.ctx, target.data_ptr(), mask.data_ptr(), None, None, self.sums.data_ptr(),
                                      self.count.data_ptr(), self.n_norm.data_ptr(), 1, 0, stream))
        for stage in range(_lib.NUM_STAGES):
            _lib.check(lib.sdn_backward_stage(eng.ctx, stage, stream))
            if self.world > 1:
                lo, hi = self.slices[stage]
                bucket = self.flat[lo:hi]
                if self.overlap:
                    self.comm_stream.wait_stream(main)
                    with torch.cuda.stream(self.comm_stream):
                        dist.all_reduce(bucket, group=self.group)
                else:
                    dist.all_reduce(bucket, group=self.group)
        if self.world > 1 and self.overlap:
            main.wait_stream(self.comm_stream)

        if isinstance(self.optimizer, FusedAdamW):
            # the "no valid pixel -> skip the step" rule is evaluated on the device: no host sync at all
            self.optimizer.step(gate=self.n_norm)
            return -1
        n_global = int(self.n_norm.cpu().item())  # the step's only host sync
        if n_global > 0 and self.optimizer is not None:
            self.optimizer.step()
        return n_global

    @torch.no_grad()
    def eval_step(self, batch: Dict[str, torch.Tensor]) -> None:
        """Validation forward + metric sums (run_epoch with optimizer=None, train.py:618)."""
        model = self.model
        lib = _lib.load()
        x = batch["input"]
        target = batch["target"].contiguous()
        mask = batch["valid_mask"].contiguous()
        self._ensure(x.device)
        model.train(False)
        model._check_input(x)
        model._launch_forward(x, True, training=False, want_outputs=False)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.sdn_loss_begin(model._engine.ctx, target.data_ptr(), mask.data_ptr(), None, None,
                                      self.sums.data_ptr(), self.count.data_ptr(), None, 0, 0, stream))


def _metrics(tot: Dict[str, float]) -> Dict[str, float]:
    n = tot["count"]
    return {
        "loss": tot["nll"] / n,
        "nll": tot["nll"] / n,
        "mae": tot["abs"] / n,
        "rmse": math.sqrt(tot["sq"] / n),
        "sigma": tot["sigma"] / n,
    }


def run_epoch(model: StereoUNet, loader: Iterable[Dict[str, torch.Tensor]], device: torch.device,
              optimizer: Optional[torch.optim.Optimizer] = None, global_step: int = 0,
              log_every_batches: Optional[int] = None, log_metrics=None,
              fused: Optional[FusedStep] = None) -> Tuple[Dict[str, float], int]:
    """Drop-in for train.py:292-418 (same arguments and return value).  ``log_metrics``
    defaults to ``mlflow.log_metrics`` when mlflow is importable; the metric names are
    the reference's (train.py:372-383)."""
    if log_metrics is None:
        try:
            import mlflow  # type: ignore

            log_metrics = mlflow.log_metrics
        except Exception:
            log_metrics = None
    is_training = optimizer is not None
    step = fused if fused is not None else FusedStep(model, optimizer)
    step.optimizer = optimizer
    step.reset_metrics()
    total = {"nll": 0.0, "abs": 0.0, "sq": 0.0, "sigma": 0.0, "count": 0}

    def flush(emit: bool) -> None:
        part = step.read_metrics()
        step.reset_metrics()
        for k in total:
            total[k] += part[k]
        if emit and is_training and log_metrics is not None and part["count"] > 0:
            m = _metrics(part)
            log_metrics({f"train_{k}_step": v for k, v in m.items()}, step=global_step)

    seen = False
    for batch in loader:
        seen = True
        if is_training:
            global_step += 1
        moved = {k: batch[k].to(device, non_blocking=True) for k in ("input", "target", "valid_mask")}
        if is_training:
            step.train_step(moved)
        else:
            step.eval_step(moved)
        if is_training and log_every_batches is not None and log_every_batches > 0 \
                and global_step % log_every_batches == 0:
            flush(True)
    if seen:
        flush(True)
    if total["count"] == 0:
        raise RuntimeError("No valid target pixels found for this epoch.")
    return _metrics(total), global_step
