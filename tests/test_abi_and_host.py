"""CPU tests: the C-ABI library loads and exports what include/sdn.h declares, and the
host-side logic (module protocol, samplers, bucket plan, N>1 plumbing on gloo)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from stereo_depth_estimation_b200 import _lib

    header = open(os.path.join(ROOT, "include", "sdn.h")).read()
    declared = set(re.findall(r"\b(sdn_[a-z_0-9]+)\s*\(", header))
    lib = _lib.load()
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sdn_version() >= 100
    assert ctypes.sizeof(_lib.AugParams) == 32
    first, num = _lib.stage_param_range(0)
    ranges = sorted(_lib.stage_param_range(s) for s in range(_lib.NUM_STAGES))
    assert ranges[0][0] == 0 and sum(n for _, n in ranges) == _lib.NUM_PARAMS
    for (f0, n0), (f1, _) in zip(ranges, ranges[1:]):
        assert f0 + n0 == f1
    assert (first, num) == (38, 28)


def test_library_fails_loudly_without_gpu():
    from stereo_depth_estimation_b200 import StereoUNet, _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    rc = _lib.load().sdn_create(ctypes.byref(ctx), 0, 1, 32, 48, 0)
    assert rc != 0 and _lib.load().sdn_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        StereoUNet()(torch.zeros(1, 6, 32, 48))


def test_module_protocol_matches_reference_layout():
    """state_dict: 120 entries, 66 parameters, 7,763,938 elements."""
    from oracle import stereo_oracle as so
    from stereo_depth_estimation_b200 import StereoUNet, load_state_dict_compat

    torch.manual_seed(42)
    model = StereoUNet(in_channels=6, out_channels=1)
    sd = model.state_dict()
    assert len(sd) == 120 and len(list(model.parameters())) == 66
    assert sum(p.numel() for p in model.parameters()) == 7_763_938
    ref = so.init_state_dict(42)
    assert [k for k in sd if so.is_param_key(k)] == so.param_keys(ref)
    for k, v in ref.items():
        assert torch.equal(sd[k], v), k          # same init under the same seed
    assert sd["enc1.block.1.num_batches_tracked"].dtype == torch.int64
    assert sd["up4.weight"].shape == (512, 256, 2, 2) and sd["dec4.block.0.weight"].shape == (256, 512, 3, 3)
    legacy = {k.replace("disparity_head", "output_head"): v for k, v in sd.items() if "logvar_head" not in k}
    other = StereoUNet()
    missing, unexpected = load_state_dict_compat(other, legacy)
    assert missing == [] and unexpected == []
    assert torch.equal(other.state_dict()["disparity_head.weight"], sd["disparity_head.weight"])
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    assert len(opt.state_dict()["param_groups"][0]["params"]) == 66
    import copy

    copy.deepcopy(model)
    assert model.training and not model.eval().training


def test_augment_sampler_ranges_and_pack():
    from stereo_depth_estimation_b200.preprocess import AugmentSampler, pack_aug

    s = AugmentSampler(seed=3)
    views = s.sample_batch(500)
    assert len(views) == 1000
    b = np.array([v.brightness for v in views])
    h = np.array([v.hue for v in views])
    g = np.array([v.gamma for v in views])
    sig = np.array([v.blur_sigma for v in views])
    nz = np.array([v.noise_std for v in views])
    sat = np.array([v.saturation for v in views])
    assert 0.8 <= b.min() and b.max() <= 1.2 and 0.75 <= sat.min() and sat.max() <= 1.25
    assert -0.09 <= h.min() and h.max() <= 0.09 and 0.8 <= g.min() and g.max() <= 1.2
    assert 0.0 <= nz.min() and nz.max() <= 0.05
    blurred = sig[sig > 0]
    assert 5 <= len(blurred) <= 70 and blurred.min() >= 0.1 and blurred.max() <= 1.0   # blur_prob 0.03
    off = AugmentSampler(0, 0, 0, 0, 0, 0, 0, 0).sample_view()
    assert (off.brightness, off.contrast, off.saturation, off.hue, off.gamma, off.blur_sigma, off.noise_std) == \
        (1, 1, 1, 0, 1, 0, 0)
    with pytest.raises(ValueError):
        AugmentSampler(blur_prob=1.5)
    packed = pack_aug(views[:4])
    assert packed.dtype == torch.uint8 and packed.numel() == 4 * 32
    assert np.frombuffer(packed.numpy().tobytes(), dtype=np.float32)[0] == np.float32(views[0].brightness)


def test_stage_slices_cover_flat_gradient_buffer():
    from stereo_depth_estimation_b200 import StereoUNet
    from stereo_depth_estimation_b200.step import stage_slices

    model = StereoUNet()
    sl = stage_slices(model)
    total = sum(p.numel() for p in model.parameters())
    assert sorted(sl)[0][0] == 0 and sorted(sl)[-1][1] == total
    assert sum(hi - lo for lo, hi in sl) == total
    # backward order: decoder tail first, encoder last; bottleneck is the 14 MB bucket
    assert sl[4][0] == 0 and (sl[2][1] - sl[2][0]) == 256 * 512 * 9 + 512 * 512 * 9 + 4 * 512
    assert sl[4][1] == sl[3][0] and (sl[4][1] - sl[4][0]) * 4 < 300_000     # the exposed last bucket is the smallest


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from stereo_depth_estimation_b200 import StereoUNet
    from stereo_depth_estimation_b200.step import stage_slices

    torch.manual_seed(0)
    model = StereoUNet()
    slices = stage_slices(model)
    total = sum(p.numel() for p in model.parameters())
    # what FusedStep does per step, on CPU tensors: all-reduce the valid count, then one
    # all-reduce per backward stage over its slice of the flat gradient buffer
    n_local = torch.tensor([1000 + 10 * rank], dtype=torch.int64)
    dist.all_reduce(n_local)
    flat = torch.full((total,), float(rank + 1))
    for lo, hi in slices:
        dist.all_reduce(flat[lo:hi])
    ok = bool((flat == sum(range(1, world + 1))).all()) and int(n_local) == sum(1000 + 10 * r for r in range(world))
    shard = 256 // world  # strong-scaling shard of the global batch
    ok = ok and shard * world == 256
    out[rank] = ok
    dist.destroy_process_group()


def test_dp_plumbing_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints one JSON line with the contract's keys (tiny run)."""
    import json
    import subprocess

    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "cpu_baseline", "e2e", "config"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["steps"] == 1 and line["warmup"] == 1          # --steps / --warmup are honoured as given
    # baseline/_ref (the vendored reference) present -> its own run_epoch is what was timed
    have_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "src", "foundation_stereo_depth", "train.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")


def test_source_prefetcher_refuses_cpu_and_bad_sources():
    """Host-side contract of the source pipeline (row N3): CUDA only, raw uint8 HWC host tensors only."""
    from stereo_depth_estimation_b200.pipeline import SourcePrefetcher

    with pytest.raises(RuntimeError):
        SourcePrefetcher([], torch.device("cpu"))
    good = torch.zeros(1, 4, 4, 3, dtype=torch.uint8)
    SourcePrefetcher._check((good, good, good))
    with pytest.raises(ValueError):
        SourcePrefetcher._check((good, good))
    with pytest.raises(ValueError):
        SourcePrefetcher._check((good, good, good.float()))
    with pytest.raises(ValueError):
        SourcePrefetcher._check((good, good, torch.zeros(1, 4, 4, 1, dtype=torch.uint8)))


def test_dropin_install_patches_reference_module_names():
    """`dropin.install()` rebinds the two names the reference imports."""
    import types

    from stereo_depth_estimation_b200 import StereoUNet, dropin, load_state_dict_compat

    pkg = types.ModuleType("foundation_stereo_depth")
    mod = types.ModuleType("foundation_stereo_depth.model")
    mod.StereoUNet = object
    mod.load_state_dict_compat = object
    pkg.model = mod
    saved = {k: sys.modules.get(k) for k in ("foundation_stereo_depth", "foundation_stereo_depth.model")}
    sys.modules["foundation_stereo_depth"] = pkg
    sys.modules["foundation_stereo_depth.model"] = mod
    try:
        dropin.install()
        assert mod.StereoUNet is StereoUNet
        assert mod.load_state_dict_compat is load_state_dict_compat
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
