"""Data-parallel gradient parity on real GPUs: run under

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_dp_parity.py

Checks, at world size N (2 here):
  1. the all-reduced flat gradient is BIT-identical on every rank;
  2. it equals the single-process gradient with the same semantics (per-rank BatchNorm statistics, loss
     normalised by the GLOBAL valid count): rank 0 recomputes every shard sequentially with n_norm = n_global
     and sums the shard gradients (only the unordered fp32 wgrad atomics differ, ~1e-6);
  3. overlapped buckets (communicator stream) == buckets all-reduced on the main stream;
  4. metric sums are global (read_metrics all-reduces), the optimizer step keeps the replicas bit-identical,
     and an all-invalid batch on ONE rank still steps (global count > 0) while an all-invalid global batch
     skips the step on every rank (train.py:331-332 decided on n_global).
Prints one JSON line; exit code 0 = all checks passed.  A summary is kept under profiles/.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from stereo_depth_estimation_b200 import StereoUNet  # noqa: E402
from stereo_depth_estimation_b200.optim import FusedAdamW  # noqa: E402
from stereo_depth_estimation_b200.step import FusedStep  # noqa: E402


def make_batch(dev, b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, 6, h, w, generator=g)
    t = torch.rand(b, 1, h, w, generator=g) * 2.0
    t[:, :, : h // 4, : w // 4] = 0.0
    return {"input": x.to(dev), "target": t.to(dev), "valid_mask": (t > 0).to(dev)}


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    b, h, w = 4, 128, 160
    shards = [make_batch(dev, b, h, w, seed=100 + r) for r in range(world)]
    # unequal valid counts per rank: mean-of-means would be wrong, the global-n normaliser is not
    shards[0]["target"][:, :, h // 2:, :] = 0.0
    shards[0]["valid_mask"] = shards[0]["target"] > 0
    n_global = sum(int((s["valid_mask"] & torch.isfinite(s["target"])).sum()) for s in shards)
    out = {"world": world, "n_global": n_global}
    ok = True

    def gather_equal(t):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return all(torch.equal(parts[0], p) for p in parts[1:])

    # ---- 1 + 3: DP step, overlapped and not
    flats = {}
    for overlap in (True, False):
        torch.manual_seed(42)
        model = StereoUNet().to(dev)
        step = FusedStep(model, optimizer=None, overlap=overlap)
        n = step.train_step(shards[rank])
        torch.cuda.synchronize()
        flats[overlap] = step.flat.clone()
        same = gather_equal(step.flat)
        out[f"identical_across_ranks_overlap{int(overlap)}"] = same
        ok &= same and n == n_global
        if overlap:
            m = step.read_metrics()
            out["global_count_from_metrics"] = m["count"]
            ok &= m["count"] == n_global
            dp_metrics = m
    e_ov = rel(flats[True], flats[False])
    out["overlap_vs_mainstream_rel"] = e_ov
    ok &= e_ov < 1e-5

    # ---- 2: single-process reference with the same semantics (rank 0 does all shards sequentially)
    total = torch.zeros_like(flats[True])
    sums = {"nll": 0.0, "abs": 0.0, "sq": 0.0, "sigma": 0.0, "count": 0}
    for r in range(world):
        torch.manual_seed(42)
        solo_model = StereoUNet().to(dev)
        solo = FusedStep(solo_model, optimizer=None)
        solo.world = 1                                   # no communicator: a plain local step ...
        cnt = torch.tensor([n_global], dtype=torch.int64, device=dev)
        solo.train_step(shards[r], valid_count=cnt)      # ... seeded with 1 / n_global
        total += solo.flat
        part = solo.read_metrics(reduce=False)
        for k in sums:
            sums[k] += part[k]
    e_sum = rel(flats[True], total)
    out["allreduced_vs_sum_of_shard_grads_rel"] = e_sum
    ok &= e_sum < 1e-5
    for k in ("nll", "abs", "sq", "sigma"):
        ok &= abs(dp_metrics[k] - sums[k]) <= 1e-9 * abs(sums[k]) + 1e-6

    # ---- 4: optimizer keeps replicas identical; skip rule on the global count
    torch.manual_seed(42)
    model = StereoUNet().to(dev)
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    step = FusedStep(model, opt)
    before = torch.cat([p.detach().flatten() for p in model.parameters()]).clone()
    for it in range(3):
        batch = make_batch(dev, b, h, w, seed=200 + 10 * it + rank)
        if it == 1 and rank == 0:
            batch["valid_mask"] = torch.zeros_like(batch["valid_mask"])     # one rank without valid pixels
        step.train_step(batch)
    torch.cuda.synchronize()
    after = torch.cat([p.detach().flatten() for p in model.parameters()])
    out["replicas_identical_after_3_steps"] = gather_equal(after)
    out["params_moved"] = bool((after != before).any())
    ok &= out["replicas_identical_after_3_steps"] and out["params_moved"]
    empty = make_batch(dev, b, h, w, seed=300 + rank)
    empty["valid_mask"] = torch.zeros_like(empty["valid_mask"])
    step.train_step(empty)
    torch.cuda.synchronize()
    frozen = torch.cat([p.detach().flatten() for p in model.parameters()])
    out["global_empty_batch_skips_step"] = bool(torch.equal(frozen, after))
    ok &= out["global_empty_batch_skips_step"]

    out["ok"] = bool(ok)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out["ok_all_ranks"] = bool(flag.item())
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
