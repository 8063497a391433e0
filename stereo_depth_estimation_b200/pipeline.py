"""Host -> device source pipeline: pinned staging + double-buffered async copy.

The reference feeds the model from CPU ``DataLoader`` workers that decode, resize and augment every
sample (dataset.py:184-311).  Here the host only hands over RAW uint8 images; ``SourcePrefetcher`` moves
the next batch's 3 x [B,Hs,Ws,3] uint8 tensors to the GPU on a copy stream while the current batch
trains, and ``DevicePreprocessor`` does the rest on the device.  The train step itself issues no
copy-engine work (augmentation parameters are staged by a kernel), so the bulk copy is the only
transfer and hides completely behind a step that is longer than it::

    pre = DevicePreprocessor(device, B, (240, 320))
    for left, right, disp, done in SourcePrefetcher(host_batches, device):
        batch = pre(left, right, disp, aug=sampler.sample_packed(B), out=out, count_out=count)
        step.train_step(batch, valid_count=count)
        done()                       # the buffers may be overwritten by the copy of batch i+2
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, Iterator, Optional, Sequence, Tuple

import torch


def bind_host_to_gpu(device: torch.device) -> Optional[list]:
    """Pin this process (and therefore the pages it first-touches, pinned staging buffers included) to the CPU
    cores NVML reports as local to ``device``.  One process per GPU under ``torchrun`` otherwise floats over both
    sockets, and at 8 ranks x 149 MB of uint8 sources per step the host -> device copies cross the inter-socket
    link and stop hiding behind the step (measured: end-to-end 8-GPU efficiency 0.70).  Call it BEFORE allocating
    the host buffers.  Returns the CPU list, or None when NVML / the affinity call is unavailable (no-op)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        dev = torch.device(device)
        props = torch.cuda.get_device_properties(dev)
        bus_id = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1 and 64 * w + b < ncpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


class SourcePrefetcher:
    """Iterates over ``batches`` (an iterable of 3-tuples of host uint8 tensors ``[B,Hs,Ws,3]``: left,
    right, RGB-encoded disparity), yielding the same tensors on ``device`` plus a ``done`` callback.

    * Host tensors that are not pinned are staged through pinned buffers (one extra host copy).
    * Two device buffer sets: the copy of batch ``i+1`` runs on a private stream while batch ``i`` is
      consumed on the caller's current stream; ``done()`` records, on that stream, that batch ``i``'s
      buffers are free again.
    * Stream-ordered only: no host synchronisation besides the one a pinned staging copy needs.
    """

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device: torch.device, depth: int = 2) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SourcePrefetcher feeds a CUDA device; there is no CPU path")
        if depth < 2:
            raise ValueError("depth must be at least 2 (one buffer in use, one in flight)")
        self.batches = batches
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._dev = [None] * depth
        self._pin = [None] * depth
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._freed = [torch.cuda.Event() for _ in range(depth)]
        self._staged = [torch.cuda.Event() for _ in range(depth)]
        self.bytes_per_batch = 0

    @staticmethod
    def _check(batch: Sequence[torch.Tensor]) -> None:
        if len(batch) != 3:
            raise ValueError("a source batch is (left, right, disparity)")
        for t in batch:
            if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3 or t.is_cuda:
                raise ValueError(f"sources must be host uint8 tensors [B,Hs,Ws,3], got {t.dtype} {tuple(t.shape)} on {t.device}")

    def _issue(self, slot: int, batch: Sequence[torch.Tensor]) -> None:
        self._check(batch)
        if self._dev[slot] is None or any(d.shape != s.shape for d, s in zip(self._dev[slot], batch)):
            # (re)allocate ON the copy stream: the caching allocator then hands out blocks that are safe to
            # write from that stream (a block freed by a main-stream tensor could still have kernels pending);
            # the buffers being dropped were last read on the consumer's stream, which _freed[slot] covers
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self._freed[slot])
                self._dev[slot] = [torch.empty(t.shape, dtype=torch.uint8, device=self.device) for t in batch]
        self.bytes_per_batch = sum(t.numel() for t in batch)
        srcs = []
        for i, t in enumerate(batch):
            t = t.contiguous()
            if not t.is_pinned():
                if self._pin[slot] is None or any(p.shape != s.shape for p, s in zip(self._pin[slot], batch)):
                    self._pin[slot] = [torch.empty(s.shape, dtype=torch.uint8).pin_memory() for s in batch]
                self._staged[slot].synchronize()          # the previous copy out of this pinned buffer is done
                self._pin[slot][i].copy_(t)
                t = self._pin[slot][i]
            srcs.append(t)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._freed[slot])
            for dst, src in zip(self._dev[slot], srcs):
                dst.copy_(src, non_blocking=True)
            self._staged[slot].record(self.copy_stream)
            self._ready[slot].record(self.copy_stream)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Callable[[], None]]]:
        it = iter(self.batches)
        main = torch.cuda.current_stream(self.device)
        for ev in self._freed:
            ev.record(main)
        pending = []
        slot = 0
        for _ in range(self.depth - 1):                      # fill the pipeline
            nxt = next(it, None)
            if nxt is None:
                break
            self._issue(slot, nxt)
            pending.append(slot)
            slot = (slot + 1) % self.depth
        while pending:
            nxt = next(it, None)
            if nxt is not None:                               # batch i+1 travels while batch i is consumed
                self._issue(slot, nxt)
                pending.append(slot)
                slot = (slot + 1) % self.depth
            cur = pending.pop(0)
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self._ready[cur])
            left, right, disp = self._dev[cur]
            for t in (left, right, disp):
                t.record_stream(main)      # allocated on the copy stream, consumed on the caller's

            def done(cur=cur):
                self._freed[cur].record(torch.cuda.current_stream(self.device))

            yield left, right, disp, done
