#!/usr/bin/env python
"""bench.py - the headline benchmark of the B200-native stereo step.

Metric (BASELINE.json): train pairs/s @320x240, plus p50 single-pair inference
latency.  Workload (BASELINE.json configs[2]): data-parallel training step,
240x320, bf16 tensor-core math with fp32 accumulate, GLOBAL batch 256,
augmentations on, at 1/2/4/8 B200 (strong scaling: each rank takes 256/N pairs).

One step = device input pipeline on raw uint8 540x960 sources (decode + resize +
rescale + L/R augmentation) -> U-Net forward (train-mode BatchNorm) -> fused
heteroscedastic loss -> backward (dgrad + wgrad + BN) -> bucketed NCCL
all-reduce (inside the library, overlapped with the backward) -> AdamW.  Nothing is
skipped or cached inside the timed region.

  value : whole-job pairs/s with the uint8 sources resident in HBM.
  e2e   : the same step fed from PINNED HOST buffers (double-buffered H2D of the
          step's sources inside the timed region) plus a D2H read of the step's
          metric sums, through the package's public API.
  roofline     : the conv_fprop family (largest tensor-core family; FIXED so that the line reports the same
                 kernel at every N), measured live with CUDA events (sdn_profile_*), algorithmic FLOPs / bytes
                 from SURVEY 8(d); `rooflines` lists every family the same way.
  dp_check     : (N > 1) the all-reduced flat gradient is bit-identical on all ranks and equals the
                 non-overlapped all-reduce of the same step.
  side_configs : (N = 1) BASELINE.json configs 4 and 5 and the torch-eager (cuDNN) bars on the same GPU,
                 the latter through the UNMODIFIED reference model from baseline/_ref.
  cpu_baseline : the reference's own run_epoch (baseline/_ref, train.py:292-418) on the host cores.
  --impl reference : the same, K steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H, W = 240, 320
HS, WS = 540, 960
GLOBAL_BATCH = 256
TRAIN_FLOPS_PER_PAIR = 85.024e9   # SURVEY 8(d) / BASELINE.md section 3
FWD_FLOPS = {(240, 320): 28.430e9, (480, 640): 113.718e9, (720, 1280): 341.154e9}
PRE_BYTES_PER_SAMPLE = 6_892_800
METRIC = "train_pairs_per_s_240x320"
UNIT = "pairs/s"
ROOFLINE_FAMILY = "conv_fprop"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--no-latency", action="store_true", help="skip the single-pair latency probe")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true", help="skip configs 4/5 and the torch-eager bars")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the step eagerly instead of "
                    "replaying the captured CUDA graph of the whole step")
    ap.add_argument("--profile-out", default="", help="write the per-op table (JSON) here")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------- CPU arm
def synth_batch_cpu(b: int, seed: int = 42):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 6, H, W, generator=g)
    t = torch.rand(b, 1, H, W, generator=g) * 64.0
    t[:, :, : H // 4, : W // 4] = 0.0
    return {"input": x, "target": t, "valid_mask": t > 0.0}


def load_reference():
    """The vendored, unmodified reference (baseline/_ref) with a no-op mlflow; None when it is absent."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import refenv

        if not refenv.available():
            return None
        train, model_mod, _ = refenv.load(fresh=True)
        return train, model_mod
    except Exception:
        return None


def cpu_train_steps(steps: int, warmup: int, batch: int = 8):
    """K timed steps of the reference loop body on the host cores (config 1: batch 8, 6x240x320, fp32,
    --device cpu) after W untimed ones.  Preferred: the reference's OWN run_epoch (train.py:292-418) on its
    own StereoUNet and torch AdamW from baseline/_ref (kind "reference"); when that copy is absent, the oracle
    port of the same loop (kind "port").  Returns (pairs/s, ms/step, threads, kind)."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    data = synth_batch_cpu(batch)
    ref = load_reference()
    if ref is not None:
        train, model_mod = ref
        torch.manual_seed(42)
        model = model_mod.StereoUNet(in_channels=6, out_channels=1)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)    # train.py:578
        dev = torch.device("cpu")
        if warmup > 0:
            train.run_epoch(model, [data] * warmup, dev, optimizer=opt, global_step=0, log_every_batches=None)
        t0 = time.perf_counter()
        train.run_epoch(model, [data] * steps, dev, optimizer=opt, global_step=0, log_every_batches=None)
        dt = (time.perf_counter() - t0) / max(steps, 1)
        return batch / dt, dt * 1e3, torch.get_num_threads(), "reference"
    from oracle import stereo_oracle as so

    sd = so.init_state_dict(42)
    opt = so.AdamWState()
    for _ in range(warmup):
        so.train_step(sd, data, opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        so.train_step(sd, data, opt)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt * 1e3, torch.get_num_threads(), "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    pps, ms, threads, kind = cpu_train_steps(steps, warm, 8)
    what = "the reference's own run_epoch (baseline/_ref, train.py:292-418) on its StereoUNet + torch AdamW" \
        if kind == "reference" else "oracle port of train.py:320-357 (fwd+bwd+loss+AdamW)"
    sample = f"{what}, fp32, batch 8 of 6x240x320 per step (BASELINE.json config 1), {threads} threads, input " \
             "pipeline excluded (the reference overlaps it in DataLoader workers)"
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C1: U-Net stereo train step, 6x240x320, batch 8, host CPU (reference --device cpu path)"},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every ~5 ms, in a thread;
    falls back to an `nvidia-smi -lms` child process when pynvml is missing)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.rows, self.proc, self.nvml, self.samples, self.stop_flag = [], None, None, [], False
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, mask))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            sm = [m for m, _ in self.samples]
            reasons = set()
            for _, mask in self.samples:
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        reasons.add(name)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ----------------------------------------------------------------- GPU arm
def synth_sources(b: int, seed: int, pinned: bool, hs: int = HS, ws: int = WS):
    """uint8 HWC sources; disparity R channel in [0,3] and ~10 % invalid."""
    rng = np.random.default_rng(seed)
    L = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
    R = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
    D = rng.integers(0, 256, (b, hs, ws, 3), dtype=np.uint8)
    D[..., 0] = rng.integers(0, 4, (b, hs, ws), dtype=np.uint8)
    D[rng.random((b, hs, ws)) < 0.1] = 0
    out = [torch.from_numpy(a) for a in (L, R, D)]
    if pinned:
        out = [t.pin_memory() for t in out]
    return out


def timed_ms(fn, warm: int, iters: int) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def family_rooflines(fam: dict, peaks: dict) -> list:
    """One entry per kernel family of the train step, same arithmetic as `roofline`."""
    out = []
    for name, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        if v["ms"] <= 0:
            continue
        if v["flops"] > 0:
            ach = v["flops"] / (v["ms"] * 1e-3) / 1e12
            out.append({"kernel": name, "bound": "tensor", "ms_per_step": v["ms"], "launches_per_step": v["calls"],
                        "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["bf16_tflops_sustained"],
                        "hbm_gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9, "hbm_frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"]})
        elif v["bytes"] > 0:
            ach = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            out.append({"kernel": name, "bound": "hbm", "ms_per_step": v["ms"], "launches_per_step": v["calls"],
                        "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]})
    return out


def side_configs(dev, peaks) -> dict:
    """BASELINE.json configs 4 and 5 plus the torch-eager (cuDNN) bars on this GPU (rank 0, N = 1)."""
    from stereo_depth_estimation_b200 import StereoUNet
    from stereo_depth_estimation_b200.preprocess import AugmentSampler, DevicePreprocessor
    from stereo_depth_estimation_b200.step import FusedStep

    out = {}
    # ---- C4: input pipeline only, 540x960 -> 240x320, batch 512 (2.39 GB of sources: larger than L2)
    B = 512
    src = [t.to(dev) for t in synth_sources(B, 7, pinned=False)]
    pre = DevicePreprocessor(dev, B, (H, W))
    sampler = AugmentSampler(seed=0)
    buf = {"o": None}

    def run_pre(aug):
        buf["o"] = pre(src[0], src[1], src[2], aug=sampler.sample_packed(B) if aug else None, out=buf["o"])

    for aug in (False, True):
        ms = timed_ms(lambda: run_pre(aug), 3, 10)
        gbs = B * PRE_BYTES_PER_SAMPLE / ms / 1e6
        out["C4_input_pipeline_b512_" + ("aug" if aug else "noaug")] = {
            "ms_per_batch": ms, "samples_per_s": B / ms * 1e3, "algorithmic_GBps": gbs, "hbm_frac": gbs / peaks["hbm_gbs"]}
    del src, buf
    pre.close()
    torch.cuda.empty_cache()
    # ---- C5: batched inference at scaled resolutions (batch 32)
    torch.manual_seed(0)
    model = StereoUNet().to(dev).eval()
    for (h, w) in ((240, 320), (480, 640), (720, 1280)):
        x = torch.rand(32, 6, h, w, device=dev)
        with torch.inference_mode():
            ms = timed_ms(lambda: model(x, return_uncertainty=True), 3, 10)
        tf = 32 * FWD_FLOPS[(h, w)] / ms / 1e9
        out[f"C5_infer_b32_{h}x{w}"] = {"ms_per_batch": ms, "pairs_per_s": 32 / ms * 1e3, "tflops": tf,
                                        "frac_of_sustained_bf16": tf / peaks["bf16_tflops_sustained"]}
        del x
        model._engine.close()
        torch.cuda.empty_cache()
    # ---- this path at batch 64 on a pre-assembled batch (like-for-like with the eager bars below)
    torch.manual_seed(0)
    model = StereoUNet().to(dev)
    step = FusedStep(model, torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4))
    g = torch.Generator().manual_seed(1)
    x = torch.rand(64, 6, H, W, generator=g).to(dev)
    t = (torch.rand(64, 1, H, W, generator=g) * 2).to(dev)
    batch = {"input": x, "target": t, "valid_mask": t > 0.2}
    ms = timed_ms(lambda: step.train_step(batch), 3, 10)
    out["b200_train_b64_preassembled"] = {"ms_per_step": ms, "pairs_per_s": 64 / ms * 1e3}
    model._engine.close()
    del model, step
    torch.cuda.empty_cache()
    # ---- torch-eager bars: the UNMODIFIED reference model + loss (baseline/_ref) on this GPU
    ref = load_reference()
    if ref is None:
        out["torch_eager"] = {"unavailable": "baseline/_ref not vendored"}
        return out
    train, model_mod = ref
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(42)
        rmodel = model_mod.StereoUNet(in_channels=6, out_channels=1).to(dev)
        ropt = torch.optim.AdamW(rmodel.parameters(), lr=1e-3, weight_decay=1e-4)
        mask = batch["valid_mask"]
        for autocast in (False, True):
            def eager_step():
                # loop body of train.py:325-343
                ropt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    pred, lv = rmodel(x, return_uncertainty=True)
                diff = pred.float()[mask] - t[mask]
                mlv = lv.float()[mask]
                loss = (diff.abs() * torch.exp(-mlv) + mlv).mean()
                loss.backward()
                ropt.step()

            rmodel.train()
            ms = timed_ms(eager_step, 2, 5)
            out["torch_eager_train_b64_" + ("bf16_autocast" if autocast else "fp32")] = {
                "ms_per_step": ms, "pairs_per_s": 64 / ms * 1e3}
            rmodel.eval()
            x1 = x[:1]

            def eager_infer():
                with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    rmodel(x1, return_uncertainty=True)

            ms = timed_ms(eager_infer, 10, 100)
            out["torch_eager_infer_1pair_" + ("bf16_autocast" if autocast else "fp32")] = {"ms": ms}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    return out


def run_b200(args):
    import torch.distributed as dist

    from stereo_depth_estimation_b200 import StereoUNet
    from stereo_depth_estimation_b200.preprocess import AugmentSampler, DevicePreprocessor
    from stereo_depth_estimation_b200.step import FusedStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from stereo_depth_estimation_b200.pipeline import bind_host_to_gpu

    numa_cpus = bind_host_to_gpu(dev) if world > 1 else None     # before any pinned allocation (first touch)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    b_local = args.global_batch // world
    if b_local * world != args.global_batch:
        raise SystemExit("global batch must divide by the number of GPUs")
    warmup = max(args.warmup, 3)

    torch.manual_seed(42)
    model = StereoUNet().to(dev)
    from stereo_depth_estimation_b200.optim import FusedAdamW

    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)  # train.py:578 hyper-parameters
    step = FusedStep(model, opt)
    pre = DevicePreprocessor(dev, b_local, (H, W))
    sampler = AugmentSampler(seed=rank)
    srcs_host = synth_sources(b_local, 1234 + rank, pinned=True)
    srcs_dev = [t.to(dev) for t in srcs_host]
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    out = None

    def eager_step(src):
        nonlocal out
        out = pre(src[0], src[1], src[2], aug=sampler.sample_packed(b_local), out=out, count_out=count)
        return step.train_step(out, valid_count=count)

    # the public whole-step API: one CUDA graph per source-buffer set (first call eager, second captures)
    from stereo_depth_estimation_b200.step import GraphedTrainStep

    gstep = None if args.no_graph else GraphedTrainStep(step, pre)

    def one_step(src):
        if gstep is None:
            return eager_step(src)
        return gstep(src[0], src[1], src[2], sampler.sample_packed(b_local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------
    l0 = model.launch_count() + pre.launch_count()
    eager_step(srcs_dev)
    launches_per_step = model.launch_count() + pre.launch_count() - l0 + 2     # + the two AdamW kernels (own context)
    for _ in range(warmup if gstep is None else max(warmup, gstep.warmup_calls)):
        one_step(srcs_dev)
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step(srcs_dev)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clock_info = clocks.stop() if clocks is not None else None
    launches = launches_per_step * args.steps      # kernels executed (graph replays re-run the captured launches)
    ms_step = ms_total / args.steps
    value = args.global_batch / (ms_step * 1e-3)

    # ---- end to end: pinned host sources -> SourcePrefetcher (the package's own pipeline API) ----
    # One continuous pipeline: warm-up steps fill it, then K steps are timed in steady state.  Every
    # timed step overlaps exactly one H2D copy (the next step's sources, issued by the prefetcher on its
    # copy stream) and ends with a D2H snapshot of its metric sums into pinned memory, which the host
    # reads (blocking) one step later so that the read never drains the launch queue.
    from stereo_depth_estimation_b200.pipeline import SourcePrefetcher

    h2d_bytes = sum(t.numel() for t in srcs_host)
    snaps = [torch.zeros(5, dtype=torch.float64).pin_memory() for _ in range(2)]
    snap_ev = [torch.cuda.Event() for _ in range(2)]
    d2h_bytes = 5 * 8

    # the bulk copy alone (no overlap), for the record
    tmp = [torch.empty_like(t, device=dev) for t in srcs_host]
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    c0.record()
    for dst, src in zip(tmp, srcs_host):
        dst.copy_(src, non_blocking=True)
    c1.record()
    torch.cuda.synchronize(dev)
    h2d_copy_ms = c0.elapsed_time(c1)
    del tmp

    def e2e_pipeline(n_warm, n_timed):
        main = torch.cuda.current_stream(dev)
        feed = SourcePrefetcher((srcs_host for _ in range(n_warm + n_timed + 1)), dev)
        seen = 0.0
        for i, (left, right, disp_src, done) in enumerate(feed):
            if i == n_warm + n_timed:
                done()
                break
            slot = i & 1
            if i == n_warm:
                e0.record(main)
            one_step((left, right, disp_src))
            done()
            packed = torch.cat([step.sums, step.count.double()])
            snaps[slot].copy_(packed, non_blocking=True)
            snap_ev[slot].record(main)
            if i > 0:                          # host-side read of the previous step's result
                snap_ev[slot ^ 1].synchronize()
                seen = float(snaps[slot ^ 1][0])
        e1.record(main)
        snap_ev[(n_warm + n_timed - 1) & 1].synchronize()
        return seen

    barrier()
    e2e_pipeline(2 if gstep is None else 2 * gstep.warmup_calls, args.steps)   # (two source slots to warm)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = args.global_batch / (e2e_ms * 1e-3)

    # ---- DP check (N > 1): all-reduced gradients identical on all ranks; overlapped == main-stream ----
    dp_check = None
    if world > 1:
        def grad_of_step(overlap: bool) -> torch.Tensor:
            step.overlap = overlap
            saved_opt, step.optimizer = step.optimizer, None     # gradients only: identical weights for both legs
            try:
                pre(srcs_dev[0], srcs_dev[1], srcs_dev[2], aug=fixed_aug, out=out, count_out=count)
                step.train_step(out, valid_count=count)
            finally:
                step.optimizer = saved_opt
            torch.cuda.synchronize(dev)
            return step.flat.clone()

        fixed_aug = sampler.sample_packed(b_local)
        g_ov = grad_of_step(True)
        g_main = grad_of_step(False)
        step.overlap = True
        digest = torch.stack([g_ov.double().sum(), g_ov.double().abs().sum(), g_ov.double().pow(2).sum()])
        parts = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(parts, digest)
        first = [torch.empty_like(g_ov[:4096]) for _ in range(world)]
        dist.all_gather(first, g_ov[:4096].contiguous())
        diff = float((g_ov.double() - g_main.double()).norm() / g_main.double().norm().clamp(min=1e-30))
        dp_check = {
            "grad_digest_identical_on_all_ranks": bool(all(torch.equal(parts[0], p) for p in parts[1:])
                                                       and all(torch.equal(first[0], f) for f in first[1:])),
            "overlapped_vs_mainstream_allreduce_rel": diff,
            "ok": bool(all(torch.equal(parts[0], p) for p in parts[1:]) and diff < 1e-5),
            "note": "same step twice (fixed augmentation parameters, no optimizer step); the only run-to-run "
                    "difference is the order of the fp32 wgrad atomics (~1e-7); see tests/gpu_dp_parity.py for the "
                    "comparison with the single-process gradient",
        }

    # ---- per-op roofline (separate, profiled steps; CUDA events per op) -----
    peaks = load_peaks()
    for _ in range(2):
        eager_step(srcs_dev)      # back on the eager launch path (the timed loops replayed graphs): settle it first
    torch.cuda.synchronize(dev)
    model.profile_enable(True)
    pre.profile_enable(True)
    prof_steps = 5
    for _ in range(prof_steps):
        eager_step(srcs_dev)     # (event-bracketed ops cannot be part of a graph)
    rows = model.profile_dump() + pre.profile_dump()
    model.profile_enable(False)
    pre.profile_enable(False)
    fam = {}
    for r in rows:
        f = fam.setdefault(r["name"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "calls": 0})
        f["ms"] += r["ms"] / prof_steps; f["flops"] += r["flops"] / prof_steps
        f["bytes"] += r["bytes"] / prof_steps; f["calls"] += r["calls"] // prof_steps
    total_prof_ms = sum(f["ms"] for f in fam.values())
    rooflines = family_rooflines(fam, peaks)
    dname = ROOFLINE_FAMILY if ROOFLINE_FAMILY in fam else max(fam.items(), key=lambda kv: kv[1]["ms"])[0]
    d = fam[dname]
    achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
    roof = {"kernel": dname, "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"],
            "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None,
            "peak_source": peaks["source"] + " (sustained bf16: kernel timed inside a long step)",
            "selection": "fixed family (the 18 conv3x3 forward launches of sdn::conv_gemm_kernel): the largest "
                         "tensor-core family; every other family is in `rooflines`"}
    # DRAM traffic per launch of that kernel family from the committed ncu --set full capture (profiles/)
    for src in ("r2_ncu_full_step_b256_summary.json", "r1_ncu_full_step_b256_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", src)) as f:
                cap = json.load(f)
            key = "conv_fprop" if "conv_fprop" in cap["families"] else "forward_conv(fprop+convT_fprop)"
            fam_cap = cap["families"][key]
            if b_local == 256:
                roof["traffic"] = fam_cap["dram_bytes"] / fam_cap["launches"]
                roof["traffic_source"] = "profiles/" + src + " (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of the " \
                                         + str(fam_cap["launches"]) + " launches of family '" + key + "' in one step, per launch)"
                roof["algorithmic_bytes_per_launch"] = fam_cap["algorithmic_bytes"] / fam_cap["launches"]
            break
        except Exception:
            continue
    roof["share_of_step"] = d["ms"] / total_prof_ms if total_prof_ms > 0 else None
    roof["launches_per_step"] = d["calls"]
    families = {k: {"ms_per_step": v["ms"],
                    "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] > 0 and v["ms"] > 0 else None,
                    "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None} for k, v in fam.items()}
    if args.profile_out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        with open(args.profile_out, "w") as f:
            json.dump({"batch_per_gpu": b_local, "rows": rows, "prof_steps": prof_steps, "families": families}, f, indent=1)

    # ---- single-pair latency (live-view path, depth_live_dl.py:518-529) -----
    latency = None
    if rank == 0 and not args.no_latency:
        model.eval()
        x_host = torch.rand(1, 6, H, W).pin_memory()
        x_dev = x_host.to(dev)
        with torch.inference_mode():
            for _ in range(20):
                model(x_dev, return_uncertainty=True)
            torch.cuda.synchronize(dev)
            dev_us, e2e_us = [], []
            for _ in range(200):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                model(x_dev, return_uncertainty=True)
                b.record()
                b.synchronize()
                dev_us.append(a.elapsed_time(b) * 1e3)
            for _ in range(200):
                t0 = time.perf_counter()
                xd = x_host.to(dev, non_blocking=True)
                dsp, lv = model(xd, return_uncertainty=True)
                dsp_h, lv_h = dsp[0, 0].cpu(), lv[0, 0].cpu()
                e2e_us.append((time.perf_counter() - t0) * 1e6)
        latency = {"device_p50_us": float(np.percentile(dev_us, 50)), "device_p99_us": float(np.percentile(dev_us, 99)),
                   "e2e_p50_us": float(np.percentile(e2e_us, 50)), "e2e_p99_us": float(np.percentile(e2e_us, 99)),
                   "what": "C2: 1x6x240x320 eval forward, disparity+logvar; e2e adds 1.84 MB H2D + 2x307 KB D2H"}
        model.train()

    # ---- side configs + CPU baseline (rank 0, N = 1 only) -------------------
    side = None
    cpu = None
    if rank == 0 and world == 1:
        if not args.no_side_configs:
            # free the train-step workspace first (45 GB at batch 256): config 5 needs 67 GB of its own
            model._engine.close()
            pre.close()
            del srcs_dev
            out = None
            torch.cuda.empty_cache()
            try:
                side = side_configs(dev, peaks)
            except Exception as exc:   # a side measurement must never cost the headline line
                side = {"error": repr(exc)[:300]}
        if not args.no_cpu_baseline:
            pps, ms, threads, kind = cpu_train_steps(3, 1, 8)
            cpu = {"value": pps, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": "3 timed steps (1 warm-up) of the reference train loop body (run_epoch, train.py:292-418: "
                             f"fwd+bwd+loss+AdamW, fp32) on batch 8 of 6x240x320, {ms:.0f} ms/step"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "warmup_extra_for_graph_capture": 0 if gstep is None else max(0, gstep.warmup_calls - warmup) + 1,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C3: DP train step 240x320, global batch %d (%d/GPU), augment on, raw uint8 "
                                   "540x960 sources -> preprocess -> fwd -> loss -> bwd -> allreduce -> AdamW"
                                   % (args.global_batch, b_local),
                       "global_batch": args.global_batch, "parallelism": f"dp{world}",
                       "launch": "eager" if gstep is None else "CUDA graph of the whole step (GraphedTrainStep), one "
                                 "cudaGraphLaunch per step",
                       "l2": "inputs larger than L2 (%.2f GB of uint8 sources per rank per step)" % (h2d_bytes / 1e9),
                       "host_affinity": ("rank 0 bound to %d GPU-local cores" % len(numa_cpus)) if numa_cpus else "unbound"},
            "model_tflops": value * TRAIN_FLOPS_PER_PAIR / 1e12,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_copy_ms_alone": h2d_copy_ms, "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": d2h_bytes * world},
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": roof,
            "rooflines": rooflines,
            "kernel_families": families,
            "dp_check": dp_check,
            "cpu_baseline": cpu,
            "latency": latency,
            "side_configs": side,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
