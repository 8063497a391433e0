#!/usr/bin/env python
"""bench.py - the headline benchmark of the B200-native stereo step.

Metric (BASELINE.json): train pairs/s @320x240, plus p50 single-pair inference
latency.  Workload (BASELINE.json configs[2]): data-parallel training step,
240x320, bf16 tensor-core math with fp32 accumulate, GLOBAL batch 256,
augmentations on, at 1/2/4/8 B200 (strong scaling: each rank takes 256/N pairs).

One step = device input pipeline on raw uint8 540x960 sources (decode + resize +
rescale + L/R augmentation) -> U-Net forward (train-mode BatchNorm) -> fused
heteroscedastic loss -> backward (dgrad + wgrad + BN) -> bucketed NCCL
all-reduce -> AdamW.  Nothing is skipped or cached inside the timed region.

  value : whole-job pairs/s with the uint8 sources resident in HBM.
  e2e   : the same step fed from PINNED HOST buffers (double-buffered H2D of the
          step's sources inside the timed region) plus a D2H read of the step's
          metric sums, through the package's public API.
  roofline     : dominant kernel family, measured live with CUDA events
                 (sdn_profile_*), algorithmic FLOPs / bytes from SURVEY 8(d).
  cpu_baseline : the oracle port of the reference train step on the host cores.
  --impl reference : the reference's own CPU path (oracle port, torch CPU fp32).
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
METRIC = "train_pairs_per_s_240x320"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--no-latency", action="store_true", help="skip the single-pair latency probe")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def cpu_train_steps(steps: int, warmup: int, batch: int = 8):
    """The reference loop body (train.py:320-357 + AdamW) restated in oracle/ and run
    with every host thread torch will use.  Returns (pairs/s, ms/step, threads)."""
    from oracle import stereo_oracle as so

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = so.init_state_dict(42)
    opt = so.AdamWState()
    data = synth_batch_cpu(batch)
    for _ in range(warmup):
        so.train_step(sd, data, opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        so.train_step(sd, data, opt)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 2))
    pps, ms, threads = cpu_train_steps(steps, warm, 8)
    sample = "oracle port of train.py:320-357 (fwd+bwd+loss+AdamW), fp32, batch 8 of 6x240x320 per step, " \
             "input pipeline excluded (the reference overlaps it in DataLoader workers)"
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C1: U-Net stereo train step, 6x240x320, batch 8, host CPU (reference --device cpu path)"},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
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
def synth_sources(b: int, seed: int, pinned: bool):
    """uint8 HWC sources; disparity R channel in [0,3] and ~10 % invalid."""
    rng = np.random.default_rng(seed)
    L = rng.integers(0, 256, (b, HS, WS, 3), dtype=np.uint8)
    R = rng.integers(0, 256, (b, HS, WS, 3), dtype=np.uint8)
    D = rng.integers(0, 256, (b, HS, WS, 3), dtype=np.uint8)
    D[..., 0] = rng.integers(0, 4, (b, HS, WS), dtype=np.uint8)
    D[rng.random((b, HS, WS)) < 0.1] = 0
    out = [torch.from_numpy(a) for a in (L, R, D)]
    if pinned:
        out = [t.pin_memory() for t in out]
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    b_local = args.global_batch // world
    if b_local * world != args.global_batch:
        raise SystemExit("global batch must divide by the number of GPUs")

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

    def one_step(src):
        nonlocal out
        out = pre(src[0], src[1], src[2], aug=sampler.sample_packed(b_local), out=out, count_out=count)
        return step.train_step(out, valid_count=count)

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
    for _ in range(max(args.warmup, 3)):
        one_step(srcs_dev)
    barrier()
    launches0 = model.launch_count() + pre.launch_count()
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
    launches = model.launch_count() + pre.launch_count() - launches0
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
    torch.cuda.synchronize(dev)
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
            packed = torch.cat([step.sums.double(), step.count.double()])
            snaps[slot].copy_(packed, non_blocking=True)
            snap_ev[slot].record(main)
            if i > 0:                          # host-side read of the previous step's result
                snap_ev[slot ^ 1].synchronize()
                seen = float(snaps[slot ^ 1][0])
        e1.record(main)
        snap_ev[(n_warm + n_timed - 1) & 1].synchronize()
        return seen

    barrier()
    e2e_pipeline(2, args.steps)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = args.global_batch / (e2e_ms * 1e-3)

    # ---- per-op roofline (separate, profiled steps; CUDA events per op) -----
    peaks = load_peaks()
    model.profile_enable(True)
    pre.profile_enable(True)
    prof_steps = 3
    for _ in range(prof_steps):
        one_step(srcs_dev)
    rows = model.profile_dump() + pre.profile_dump()
    model.profile_enable(False)
    pre.profile_enable(False)
    fam = {}
    for r in rows:
        f = fam.setdefault(r["name"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "calls": 0})
        f["ms"] += r["ms"] / prof_steps; f["flops"] += r["flops"] / prof_steps
        f["bytes"] += r["bytes"] / prof_steps; f["calls"] += r["calls"] // prof_steps
    total_prof_ms = sum(f["ms"] for f in fam.values())
    dominant = max(fam.items(), key=lambda kv: kv[1]["ms"])
    dname, d = dominant
    if d["flops"] > 0:
        achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
        roof = {"kernel": dname, "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None,
                "peak_source": peaks["source"] + " (sustained bf16: kernel timed inside a long step)"}
    else:
        achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        roof = {"kernel": dname, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"]}
    # DRAM traffic per launch of that kernel family from the committed ncu --set full capture (profiles/)
    try:
        src = "r1_ncu_full_step_b256_summary.json"
        with open(os.path.join(ROOT, "profiles", src)) as f:
            cap = json.load(f)
        if dname == "bn_bwd" and b_local == 256:
            fam_cap = cap["families"]["bn_bwd"]
            roof["traffic"] = fam_cap["dram_bytes"] / fam_cap["layers_captured"]
            roof["traffic_source"] = "profiles/" + src + " (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of the" \
                                     " reduce + apply kernels, averaged over the 18 layers of one step)"
            roof["algorithmic_bytes_per_launch"] = d["bytes"] / max(d["calls"], 1)
        elif dname == "conv_fprop" and b_local == 256:
            fam_cap = cap["families"]["forward_conv(fprop+convT_fprop)"]
            roof["traffic"] = fam_cap["dram_bytes"] / fam_cap["launches"]
            roof["traffic_source"] = "profiles/" + src + " (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of the" \
                                     " 22 forward conv_gemm launches of one step = 18 conv fprop + 4 ConvTranspose2d fprop)"
            roof["algorithmic_bytes_per_launch"] = fam_cap["algorithmic_bytes"] / fam_cap["launches"]
    except Exception:
        pass
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
                   "what": "1x6x240x320 eval forward, disparity+logvar; e2e adds 1.84 MB H2D + 2x307 KB D2H"}
        model.train()

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pps, ms, threads = cpu_train_steps(3, 1, 8)
        cpu = {"value": pps, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "3 steps of the oracle train step (fwd+bwd+loss+AdamW, fp32) on batch 8 of 6x240x320, "
                         f"{ms:.0f} ms/step"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C3: DP train step 240x320, global batch %d (%d/GPU), augment on, raw uint8 "
                                   "540x960 sources -> preprocess -> fwd -> loss -> bwd -> allreduce -> AdamW"
                                   % (args.global_batch, b_local),
                       "global_batch": args.global_batch, "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (%.2f GB of uint8 sources per rank per step)" % (h2d_bytes / 1e9)},
            "model_tflops": value * TRAIN_FLOPS_PER_PAIR / 1e12,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_copy_ms_alone": h2d_copy_ms, "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": d2h_bytes * world},
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": roof,
            "kernel_families": families,
            "cpu_baseline": cpu,
            "latency": latency,
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
