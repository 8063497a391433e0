"""ctypes binding of libsdn_b200.so (the C ABI declared in include/sdn.h).

There is deliberately no fallback: if the CUDA library is missing, every entry
point of this package raises.  Build it with ``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C stereo_depth_estimation_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_uint, c_uint8, c_uint32, c_ulonglong, c_void_p

# SDN_LIB_NAME selects an alternative in-tree build (e.g. the -DSDN_FORENSICS timing-experiment variant)
LIB_NAME = os.environ.get("SDN_LIB_NAME", "libsdn_b200.so")
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

NUM_PARAMS = 66
NUM_BN = 18
NUM_STAGES = 5
RESIZE_FOURTERM = 1
PREPROCESS_AUG_HOST = 4
CTX_PREPROCESS_ONLY = 1
STEP_HAVE_COUNT = 1
STEP_NO_OVERLAP = 2
F32, F64, U64 = 0, 1, 2

# every symbol include/sdn.h declares (tests check that the .so exports them all)
EXPORTS = (
    "sdn_last_error",
    "sdn_version",
    "sdn_create",
    "sdn_destroy",
    "sdn_workspace_bytes",
    "sdn_set_params",
    "sdn_forward",
    "sdn_backward_begin",
    "sdn_backward_stage",
    "sdn_stage_param_range",
    "sdn_loss_begin",
    "sdn_count_valid",
    "sdn_train_step",
    "sdn_eval_step",
    "sdn_comm_unique_id",
    "sdn_comm_init",
    "sdn_comm_destroy",
    "sdn_comm_world",
    "sdn_comm_allreduce",
    "sdn_preprocess",
    "sdn_preprocess_cached",
    "sdn_live_preprocess",
    "sdn_live_postprocess",
    "sdn_debug_read",
    "sdn_adamw_step",
    "sdn_debug_trace",
    "sdn_profile_enable",
    "sdn_profile_dump",
    "sdn_launch_count",
)


class AugParams(ctypes.Structure):
    """sdn_aug_params (include/sdn.h): one view's photometric parameters."""

    _fields_ = [
        ("brightness", c_float),
        ("contrast", c_float),
        ("saturation", c_float),
        ("hue", c_float),
        ("gamma", c_float),
        ("blur_sigma", c_float),
        ("noise_std", c_float),
        ("noise_seed", c_uint32),
    ]


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_NAME} is not built ({LIB_PATH} missing). The B200-native stereo path has no "
            "CPU/PyTorch fallback; run __graft_entry__.build() first."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.sdn_last_error.restype = c_char_p
    lib.sdn_last_error.argtypes = []
    lib.sdn_version.restype = c_int
    lib.sdn_create.restype = c_int
    lib.sdn_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_uint]
    lib.sdn_destroy.restype = c_int
    lib.sdn_destroy.argtypes = [c_void_p]
    lib.sdn_workspace_bytes.restype = c_int64
    lib.sdn_workspace_bytes.argtypes = [c_void_p]
    lib.sdn_launch_count.restype = c_int64
    lib.sdn_launch_count.argtypes = [c_void_p]
    lib.sdn_set_params.restype = c_int
    lib.sdn_set_params.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                   POINTER(c_void_p), POINTER(c_void_p)]
    lib.sdn_forward.restype = c_int
    lib.sdn_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.sdn_backward_begin.restype = c_int
    lib.sdn_backward_begin.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    lib.sdn_backward_stage.restype = c_int
    lib.sdn_backward_stage.argtypes = [c_void_p, c_int, c_void_p]
    lib.sdn_stage_param_range.restype = c_int
    lib.sdn_stage_param_range.argtypes = [c_int, POINTER(c_int), POINTER(c_int)]
    lib.sdn_loss_begin.restype = c_int
    lib.sdn_loss_begin.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_void_p]
    lib.sdn_count_valid.restype = c_int
    lib.sdn_count_valid.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.sdn_train_step.restype = c_int
    lib.sdn_train_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_uint, c_void_p]
    lib.sdn_eval_step.restype = c_int
    lib.sdn_eval_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]
    lib.sdn_comm_unique_id.restype = c_int
    lib.sdn_comm_unique_id.argtypes = [c_void_p]
    lib.sdn_comm_init.restype = c_int
    lib.sdn_comm_init.argtypes = [c_void_p, c_void_p, c_int, c_int]
    lib.sdn_comm_destroy.restype = c_int
    lib.sdn_comm_destroy.argtypes = [c_void_p]
    lib.sdn_comm_world.restype = c_int
    lib.sdn_comm_world.argtypes = [c_void_p]
    lib.sdn_comm_allreduce.restype = c_int
    lib.sdn_comm_allreduce.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p]
    lib.sdn_preprocess.restype = c_int
    lib.sdn_preprocess.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_uint, c_void_p]
    lib.sdn_preprocess_cached.restype = c_int
    lib.sdn_preprocess_cached.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_uint, c_void_p]
    lib.sdn_live_preprocess.restype = c_int
    lib.sdn_live_preprocess.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.sdn_live_postprocess.restype = c_int
    lib.sdn_live_postprocess.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.sdn_adamw_step.restype = c_int
    lib.sdn_adamw_step.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                   POINTER(c_int64), c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, c_void_p,
                                   c_void_p, c_void_p]
    lib.sdn_debug_trace.restype = c_int
    lib.sdn_debug_trace.argtypes = [c_void_p, c_void_p]
    lib.sdn_profile_enable.restype = c_int
    lib.sdn_profile_enable.argtypes = [c_void_p, c_int]
    lib.sdn_profile_dump.restype = c_int
    lib.sdn_profile_dump.argtypes = [c_void_p, ctypes.c_char_p, c_int64]
    lib.sdn_debug_read.restype = c_int
    lib.sdn_debug_read.argtypes = [c_void_p, c_int, c_int, POINTER(c_float), c_int64, POINTER(c_int)]
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Raise the library's error message as a Python exception (the reference has
    no error codes: it raises, e.g. train.py:390-391, dataset.py:166-182)."""
    if rc != 0:
        msg = load().sdn_last_error()
        raise RuntimeError("libsdn_b200: " + (msg.decode("utf-8", "replace") if msg else f"error {rc}"))


def profile_dump(ctx) -> list:
    """Parse sdn_profile_dump's CSV into a list of dicts."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(load().sdn_profile_dump(ctx, buf, len(buf)))
    rows = []
    lines = buf.value.decode().strip().splitlines()
    for line in lines[1:]:
        name, layer, calls, ms, flops, nbytes = line.split(",")
        rows.append({"name": name, "layer": int(layer), "calls": int(calls), "ms": float(ms), "flops": float(flops),
                     "bytes": float(nbytes)})
    return rows


def stage_param_range(stage: int) -> tuple[int, int]:
    first, num = c_int(0), c_int(0)
    check(load().sdn_stage_param_range(stage, ctypes.byref(first), ctypes.byref(num)))
    return first.value, num.value


_ = (c_uint8, c_ulonglong)  # re-exported for callers building argument buffers
