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
single-process gradient exactly; the backward stages (5) are the all-reduce
buckets, issued on a side stream as soon as each stage's kernels are enqueued so
NCCL overlaps the rest of the backward.  BatchNorm statistics stay per rank
(standard DDP semantics).
"""
from __future__ import annotations

import ctypes
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
    """Level-B2 step driver over ``sdn_train_step`` / ``sdn_eval_step`` (include/sdn.h).

    ``process_group``: a ``torch.distributed`` group (any backend) used ONLY to hand rank 0's NCCL unique id
    to the other ranks; the data path (valid-count and gradient-bucket all-reduces) runs on the library's own
    communicator (``sdn_comm_init``) so that the buckets overlap the backward inside one C call."""

    def __init__(self, model: StereoUNet, optimizer: Optional[torch.optim.Optimizer] = None,
                 process_group=None, overlap: bool = True) -> None:
        self.model = model
        self.optimizer = optimizer
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist is not None and dist.is_initialized()) else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        self.overlap = overlap
        self.flat = None
        self.views = None
        self.slices = None
        self.sums = None       # device fp64 [4]: sum nll, |diff|, diff^2, exp(.5 logvar)
        self.count = None      # device i64 [1]: valid pixels accumulated with the sums
        self.n_norm = None     # device i64 [1]: loss normaliser of the current step (global count under DP)
        self._comm_ctx = None  # address of the sdn_ctx that holds the communicator

    # -------------------------------------------------------------- buffers
    def _ensure(self, device: torch.device) -> None:
        if self.flat is None or self.flat.device != device:
            self.flat, self.views = self.model._new_grad_views(device)
            self.flat.zero_()
            self.slices = stage_slices(self.model)
            self.sums = torch.zeros(4, device=device, dtype=torch.float64)
            self.count = torch.zeros(1, device=device, dtype=torch.int64)
            self.n_norm = torch.zeros(1, device=device, dtype=torch.int64)
            # the reference loop calls optimizer.zero_grad(set_to_none=True) every step (train.py:325): re-attach
        for p, v in zip(self.model._param_list(), self.views):
            if p.grad is not v:
                p.grad = v

    def _ensure_comm(self) -> None:
        """Collective: every rank calls it on its first step (and again if the context was re-created)."""
        eng = self.model._engine
        if self.world == 1 or self._comm_ctx == eng.ctx.value:
            return
        lib = _lib.load()
        box = [None]
        if self.rank == 0:
            raw = ctypes.create_string_buffer(128)
            _lib.check(lib.sdn_comm_unique_id(raw))
            box[0] = raw.raw
        dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                   group=self.group)
        _lib.check(lib.sdn_comm_init(eng.ctx, ctypes.c_char_p(box[0]), self.rank, self.world))
        self._comm_ctx = eng.ctx.value

    def reset_metrics(self) -> None:
        if self.sums is not None:
            self.sums.zero_()
            self.count.zero_()

    def read_metrics(self, reduce: bool = True) -> Dict[str, float]:
        """One D2H copy of the five running sums (train.py:345-357).  Under data parallelism the sums of all
        ranks are added first (collective: every rank must call it), so the metrics describe the global batch
        like the single-process reference's do; ``reduce=False`` returns this rank's share."""
        sums, count = self.sums, self.count
        if reduce and self.world > 1 and self._comm_ctx is not None:
            lib = _lib.load()
            eng = self.model._engine
            stream = torch.cuda.current_stream(sums.device).cuda_stream
            sums, count = sums.clone(), count.clone()
            _lib.check(lib.sdn_comm_allreduce(eng.ctx, sums.data_ptr(), 4, _lib.F64, stream))
            _lib.check(lib.sdn_comm_allreduce(eng.ctx, count.data_ptr(), 1, _lib.U64, stream))
        s = sums.cpu()
        n = int(count.cpu().item())
        return {"nll": float(s[0]), "abs": float(s[1]), "sq": float(s[2]), "sigma": float(s[3]), "count": n}

    # ----------------------------------------------------------------- steps
    def train_step(self, batch: Dict[str, torch.Tensor], valid_count: Optional[torch.Tensor] = None) -> int:
        """One optimisation step on an already-assembled batch (the reference's sample
        format).  ``valid_count`` (device int64 [1]) may come from the preprocessing
        kernel (it is all-reduced IN PLACE under data parallelism); otherwise it is counted here.
        Returns the (global) valid count; 0 means
        the step was skipped like train.py:331-332 (-1 with ``FusedAdamW``: the rule is applied on the
        device and the host never learns the count)."""
        model = self.model
        lib = _lib.load()
        x = batch["input"]
        target = batch["target"].contiguous()
        mask = batch["valid_mask"].contiguous()
        device = x.device
        model._check_input(x)
        if target.dtype != torch.float32 or mask.dtype not in (torch.bool, torch.uint8):
            raise ValueError("target must be float32 and valid_mask bool (the reference's sample format)")
        self._ensure(device)
        model.train(True)
        x, dirty = model._pre_forward(x, True)
        eng = model._engine
        if eng.grad_views is None or len(eng.grad_views) != len(self.views) or \
                any(a is not b for a, b in zip(eng.grad_views, self.views)):
            model._bind_grads(self.views)     # (the autograd path, or a new context, had bound something else)
        self._ensure_comm()
        n_norm = self.n_norm if valid_count is None else valid_count.view(1)
        if n_norm.dtype != torch.int64 or not n_norm.is_cuda:
            raise ValueError("valid_count must be a CUDA int64 tensor [1]")
        flags = (_lib.STEP_HAVE_COUNT if valid_count is not None else 0) | (0 if self.overlap else _lib.STEP_NO_OVERLAP)
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(lib.sdn_train_step(eng.ctx, x.data_ptr(), target.data_ptr(), mask.data_ptr(), x.shape[0],
                                      1 if dirty else 0, self.sums.data_ptr(), self.count.data_ptr(),
                                      n_norm.data_ptr(), flags, stream))
        model._post_forward(dirty, True)

        if isinstance(self.optimizer, FusedAdamW):
            # the "no valid pixel -> skip the step" rule is evaluated on the device: no host sync at all
            self.optimizer.step(gate=n_norm)
            return -1
        n_global = int(n_norm.cpu().item())  # the step's only host sync
        if n_global > 0 and self.optimizer is not None:
            self.optimizer.step()
        return n_global

    @torch.no_grad()
    def eval_step(self, batch: Dict[str, torch.Tensor], want_outputs: bool = False):
        """Validation forward + metric sums (run_epoch with optimizer=None, train.py:618).  With
        ``want_outputs`` also returns (disparity, logvar) like the preview path (train.py:268-272)."""
        model = self.model
        lib = _lib.load()
        x = batch["input"]
        target = batch["target"].contiguous()
        mask = batch["valid_mask"].contiguous()
        model._check_input(x)
        self._ensure(x.device)
        model.train(False)
        x, dirty = model._pre_forward(x, False)
        b, _, h, w = x.shape
        disp = torch.empty((b, 1, h, w), device=x.device, dtype=torch.float32) if want_outputs else None
        logvar = torch.empty_like(disp) if want_outputs else None
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.sdn_eval_step(model._engine.ctx, x.data_ptr(), target.data_ptr(), mask.data_ptr(), b,
                                     1 if dirty else 0, disp.data_ptr() if disp is not None else None,
                                     logvar.data_ptr() if logvar is not None else None, self.sums.data_ptr(),
                                     self.count.data_ptr(), stream))
        model._post_forward(dirty, False)
        return (disp, logvar) if want_outputs else None


class GraphedTrainStep:
    """The whole train step - device input pipeline, ``sdn_train_step`` (forward, fused loss, backward, bucketed
    all-reduce on the library's communicator) and ``FusedAdamW`` - captured ONCE per source-buffer set as a CUDA
    graph and replayed: ~210 kernel launches become one ``cudaGraphLaunch`` per step.  It pays where the per-GPU
    batch is small (32 pairs per GPU in the 8-GPU run: the step is 6 ms and the gaps between its kernels show).

    A graph bakes device addresses in, so one graph is kept per (left, right, disparity) buffer triple - the two
    slots of ``SourcePrefetcher`` give two graphs - together with its own pinned augmentation-parameter buffer, which
    the staging kernel inside the graph reads in place.  The first call for a triple runs eagerly (it also warms the
    library up), the second captures, later ones replay.  Needs ``FusedAdamW`` (the skip-empty-batch rule is on the
    device; a host-side ``optimizer.step()`` cannot be captured)."""

    def __init__(self, step: "FusedStep", pre, out: Optional[dict] = None) -> None:
        if not isinstance(step.optimizer, FusedAdamW):
            raise ValueError("GraphedTrainStep needs FusedAdamW (the optimizer step is part of the graph)")
        self.step, self.pre = step, pre
        self.out = out
        self.count = None
        self.entries = {}      # (ptrs, parity) -> dict(graph, aug, event, calls)
        self.turns = {}

    def __call__(self, left: torch.Tensor, right: torch.Tensor, disparity: torch.Tensor, aug: torch.Tensor) -> None:
        """``aug``: host uint8 tensor from ``AugmentSampler.sample_packed`` (2*B parameter sets)."""
        dev = left.device
        if self.count is None:
            self.count = torch.zeros(1, dtype=torch.int64, device=dev)
        base = (left.data_ptr(), right.data_ptr(), disparity.data_ptr(), tuple(left.shape))
        # two graphs (two pinned parameter buffers) per buffer triple, used alternately: the host may prepare step
        # k+1 while step k still runs, it only waits for the replay of two steps ago to have read its parameters
        turn = self.turns.get(base, 0)
        self.turns[base] = turn + 1
        key = base + (turn & 1,)
        ent = self.entries.get(key)
        if ent is None:
            ent = {"graph": None, "aug": torch.empty(aug.numel(), dtype=torch.uint8).pin_memory(),
                   "event": torch.cuda.Event(), "calls": 0}
            self.entries[key] = ent
        ent["event"].synchronize()           # the previous replay of THIS graph has read its parameter buffer
        ent["aug"].copy_(aug)
        stream = torch.cuda.current_stream(dev)

        def body():
            self.out = self.pre(left, right, disparity, aug=ent["aug"], out=self.out, count_out=self.count)
            self.step.train_step(self.out, valid_count=self.count)

        if ent["graph"] is None and ent["calls"] >= 1:
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            ent["graph"] = graph
        if ent["graph"] is not None:
            ent["graph"].replay()
            # the graph updated the parameters behind Python's back: the packed operand cache is stale for
            # whoever calls the module next (the captured step re-packs on every replay by construction)
            eng = self.step.model._engine
            eng.packed_versions = None
            eng.generation += 1
        else:
            body()
        ent["calls"] += 1
        ent["event"].record(stream)

    @property
    def warmup_calls(self) -> int:
        """Calls per buffer triple before every step is a replay (2 eager + 2 capturing)."""
        return 4


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
