"""ctypes bridge to the sm_100a device library (the C ABI of include/nanograd_b200.h).

There is no CPU fallback: if the shared library has not been built, or no B200 is visible,
the first device operation raises.  Build with ``python -m nanograd_b200.build`` (or
``__graft_entry__.build()``); the library lives in-tree at ``nanograd_b200/lib``.

Replaces the PyOpenCL plumbing of the reference (nanograd/tensor.py:14-30,167-180 and the
``cl.Program(...).build()`` calls scattered through nanograd/nn/ops_gpu.py).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_PKG, "lib", "libnanograd_b200.so")

_lib = None
_lib_path = None

u64, i64, i32, f32 = c_uint64, c_int64, c_int, c_float
_byref = ctypes.byref
p_i64, p_u64, p_i32 = POINTER(c_int64), POINTER(c_uint64), POINTER(c_int32)

# name -> argtypes ; every function returns int status except ng_last_error
_PROTOTYPES = {
    "ng_init": [i32],
    "ng_device_info": [POINTER(c_int), POINTER(c_int), POINTER(c_int), p_u64],
    "ng_sync": [],
    "ng_launch_count": [p_u64],
    "ng_malloc": [p_u64, u64],
    "ng_free": [u64],
    "ng_mem_stats": [p_u64, p_u64, p_u64],
    "ng_pool_release": [],
    "ng_memset_zero": [u64, u64],
    "ng_h2d": [u64, c_void_p, u64],
    "ng_d2h": [c_void_p, u64, u64],
    "ng_d2h_after": [c_void_p, u64, u64, u64],
    "ng_d2d": [u64, u64, u64],
    "ng_h2d_prefetch": [i32, u64, c_void_p, u64],
    "ng_prefetch_wait": [i32],
    "ng_prefetch_sync": [i32],
    "ng_host_gather_rows": [c_void_p, c_void_p, p_i64, i64, u64, i64],
    "ng_loader_create": [p_u64, c_void_p, c_void_p, i64, u64, i32, i32, p_i64, p_i64, u64, u64, u64, u64],
    "ng_loader_next": [u64, POINTER(c_int), p_i64],
    "ng_loader_destroy": [u64],
    "ng_argmax_rows": [u64, u64, i64, i64],
    "ng_count_equal": [u64, u64, u64, i64],
    "ng_pinned_alloc": [POINTER(c_void_p), u64],
    "ng_pinned_free": [c_void_p],
    "ng_event_create": [p_u64],
    "ng_event_record": [u64],
    "ng_event_sync": [u64],
    "ng_event_elapsed_ms": [u64, u64, POINTER(c_float)],
    "ng_event_destroy": [u64],
    "ng_flush_l2": [],
    "ng_measure_ffma_peak": [i32, POINTER(c_float)],
    "ng_prof_enable": [i32],
    "ng_prof_reset": [],
    "ng_prof_read": [i32, POINTER(ctypes.c_double), p_i64, POINTER(ctypes.c_double), POINTER(ctypes.c_double)],
    "ng_unary": [i32, u64, u64, i64],
    "ng_binary": [i32, u64, u64, u64, i32, p_i64, p_i64, p_i64],
    "ng_binary_scalar": [i32, u64, u64, f32, i32, i64],
    "ng_fill": [u64, f32, i64],
    "ng_accumulate": [u64, u64, i64],
    "ng_reduce": [i32, u64, u64, i32, p_i64, p_i32],
    "ng_channel_sum": [u64, u64, i64, i64, i64],
    "ng_slice": [u64, u64, i32, p_i64, p_i64, p_i64],
    "ng_permute": [u64, u64, i32, p_i64, p_i32],
    "ng_onehot": [u64, u64, i64, i64],
    "ng_gemm": [u64, u64, u64, i64, i64, i64, i32, i32, i64, u64, i32],
    "ng_conv2d_fprop": [u64, u64, u64, u64] + [i32] * 13,
    "ng_conv2d_dgrad": [u64, u64, u64] + [i32] * 13,
    "ng_conv2d_wgrad": [u64, u64, u64] + [i32] * 13,
    "ng_conv2d_tc_plan": [i32] * 11 + [POINTER(c_int)],
    "ng_conv2d_stage_nhwc": [u64, u64, i64, i64, i64, i32],
    "ng_conv2d_stage_nhwc_sum": [u64, u64, u64, i64, i64, i64, i32],
    "ng_maxpool2d_fwd_bn": [u64] * 6 + [i64, i32, i32, i32, i32, i32, i32],
    "ng_maxpool2d_bwd_bn": [u64] * 8 + [i64, i32, i32, i32, i32, i32, i32],
    "ng_conv2d_stage_f16": [u64, u64, u64, u64, i64, i64, i64],
    "ng_conv2d_stage_filters_f16": [u64, u64, u64, i32, i32, i32],
    "ng_conv2d_stage_f16_bn": [u64] * 7 + [i64, i64, i64, i32],
    "ng_conv2d_stage_f16_bnbwd": [u64] * 10 + [i64, i64, i64, i32],
    "ng_bn2d_act_fwd_stats": [u64] * 9 + [i64, i64, i64, f32, f32, i32],
    "ng_bn2d_act_apply": [u64] * 6 + [i64, i64, i64, i32],
    "ng_bn2d_act_bwd_sums": [u64] * 10 + [i64, i64, i64, i32],
    "ng_bn2d_act_bwd_dx": [u64] * 8 + [i64, i64, i64, i32],
    "ng_bn2d_act_bwd_dx_sum": [u64] * 9 + [i64, i64, i64, i32],
    "ng_bn2d_act_bwd_sums_pool": [u64] * 11 + [i64, i64, i32, i32, i32],
    "ng_conv2d_stage_f16_bnbwd_pool": [u64] * 11 + [i64, i64, i32, i32, i32],
    "ng_maxpool2d_fwd": [u64, u64, i64, i32, i32, i32, i32],
    "ng_maxpool2d_bwd": [u64, u64, u64, u64, i64, i32, i32, i32, i32],
    "ng_avgpool2d_fwd": [u64, u64, i64, i32, i32, i32, i32],
    "ng_avgpool2d_bwd": [u64, u64, i64, i32, i32, i32, i32],
    "ng_logsoftmax_fwd": [u64, u64, i64, i64],
    "ng_logsoftmax_bwd": [u64, u64, u64, i64, i64],
    "ng_nll_fwd": [u64, u64, u64, i64, i64],
    "ng_nll_bwd": [u64, u64, u64, i64, i64],
    "ng_mse_fwd": [u64, u64, u64, i64],
    "ng_mse_bwd": [u64, u64, u64, u64, i64, i32],
    "ng_act_fwd": [i32, u64, u64, u64, f32, i64],
    "ng_act_bwd": [i32, u64, u64, u64, u64, f32, i64],
    "ng_swish_dbeta": [u64, u64, u64, u64, i64],
    "ng_bn2d_fwd_train": [u64] * 8 + [i64, i64, i64, f32, f32],
    "ng_bn2d_fwd_eval": [u64] * 6 + [i64, i64, i64, f32],
    "ng_bn2d_bwd": [u64] * 8 + [i64, i64, i64],
    "ng_bn2d_act_fwd_train": [u64] * 8 + [i64, i64, i64, f32, f32, i32],
    "ng_bn2d_act_fwd_eval": [u64] * 6 + [i64, i64, i64, f32, i32],
    "ng_bn2d_act_bwd": [u64] * 9 + [i64, i64, i64, i32],
    "ng_multi_sgd": [i32, p_u64, p_u64, p_u64, p_i64, f32, f32, f32],
    "ng_multi_adam": [i32, p_u64, p_u64, p_u64, p_u64, p_i64, f32, f32, f32, f32, f32, f32, f32, i32, f32],
    "ng_multi_pack": [i32, p_u64, p_i64, u64],
    "ng_multi_unpack": [i32, p_u64, p_i64, u64],
    "ng_graph_begin": [],
    "ng_graph_end": [p_u64, p_u64, p_u64],
    "ng_graph_launch": [u64],
    "ng_graph_destroy": [u64],
    "ng_graph_capturing": [POINTER(c_int)],
    "ng_nccl_unique_id": [c_void_p],
    "ng_nccl_init": [c_void_p, i32, i32],
    "ng_nccl_world": [POINTER(c_int), POINTER(c_int)],
    "ng_nccl_allreduce_sum": [u64, i64],
    "ng_nccl_allreduce_begin": [u64, i64],
    "ng_nccl_allreduce_end": [],
    "ng_nccl_sync_bn": [i32],
    "ng_nccl_host_allreduce": [POINTER(ctypes.c_double), i32, i32],
    "ng_nccl_destroy": [],
}

EXPORTED_SYMBOLS = sorted(list(_PROTOTYPES) + ["ng_last_error"])

# op codes (mirrors the enums of include/nanograd_b200.h)
U_COPY, U_NEG, U_EXP, U_LOG, U_RELU, U_SIGMOID, U_TANH, U_SQRT, U_RECIP = range(9)
ACT_LEAKY_RELU, ACT_MISH, ACT_SWISH = range(3)
(B_ADD, B_MUL, B_POW, B_DIV, B_SUB, B_EQ, B_RELU_BWD, B_EXP_BWD, B_LOG_BWD, B_SIGMOID_BWD,
 B_TANH_BWD, B_POW_DBASE, B_POW_DEXP) = range(13)
R_SUM, R_MAX, R_MIN = range(3)
F_RELU, F_TF32, F_TF32X3 = 1, 2, 4

# Math mode of the convolution kernels (fprop, dgrad, wgrad):
#   "fp32" (default)  FP32-accurate results (parity bar rtol 1e-4).  Per layer the library picks the
#                     fastest FP32-accurate engine: for problems large enough to pay for staging, the
#                     tcgen05 tensor-core kernels — error-compensated FP16x3 operand pairs (channel counts
#                     multiples of 64; measured 1e-6 .. 8e-6 from the SIMT kernels) or the 3xTF32 operand
#                     split (other supported shapes; 2e-6 .. 1.4e-5 from a float64 reference) — else the
#                     FP32 SIMT (FFMA) kernel.
#   "simt"            FP32 SIMT kernels only.
#   "tf32x3"          3xTF32 tensor-core kernels for every supported shape, whatever its size.
#   "f16x3"           FP16x3 tensor-core kernels for every supported shape, 3xTF32 for the others.
#   "fp32_tf32x3"     the default selection without the FP16x3 engine (round-1 behaviour).
#   "tf32"  (opt-in)  tensor cores with plain TF32 inputs / FP32 accumulation (parity bar rtol 1e-2).
# Selected with NANOGRAD_B200_MATH (or the reference-survey spelling NANOGRAD_TF32=1) or set_math_mode().
F_TC_FORCE, F_X_NHWC, F_DY_NHWC, F_F16X3, F_STAGE_F16, F_W_STAGED = 8, 16, 32, 64, 128, 256
_MATH_MODES = {"fp32": F_F16X3 | F_TF32X3, "simt": 0, "tf32x3": F_TF32X3 | F_TC_FORCE, "tf32": F_TF32,
               "fp32_tf32x3": F_TF32X3, "f16x3": F_F16X3 | F_TF32X3 | F_TC_FORCE}
_math_flags = _MATH_MODES["tf32" if os.environ.get("NANOGRAD_TF32", "0") == "1"
                          else os.environ.get("NANOGRAD_B200_MATH", "fp32").lower()]


def set_math_mode(mode):
    global _math_flags
    if mode not in _MATH_MODES:
        raise ValueError("math mode must be one of %s" % sorted(_MATH_MODES))
    _math_flags = _MATH_MODES[mode]


def math_mode():
    return [k for k, v in _MATH_MODES.items() if v == _math_flags][0]


def math_flags():
    return _math_flags


class DeviceLibraryError(Exception):
    """Raised when the device library is missing or a device call fails."""


def _errcheck(result, func, args):
    if result != 0:
        msg = _lib.ng_last_error()
        raise DeviceLibraryError("%s: %s" % (func.__name__, msg.decode("utf-8", "replace") if msg else "failed"))
    return result


def load_library(path=None):
    """Load the C-ABI shared library and bind every prototype.  Returns the ctypes handle."""
    global _lib, _lib_path
    path = path or os.environ.get("NANOGRAD_B200_LIB") or DEFAULT_LIB
    if _lib is not None and _lib_path == path:
        return _lib
    if not os.path.exists(path):
        raise DeviceLibraryError(
            "nanograd_b200 device library not found at %s. Build it with `python -m nanograd_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for Device.GPU." % path)
    handle = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    handle.ng_last_error.restype = c_char_p
    handle.ng_last_error.argtypes = []
    for name, argtypes in _PROTOTYPES.items():
        fn = getattr(handle, name)  # AttributeError here = ABI mismatch, fail loudly
        fn.argtypes = argtypes
        fn.restype = c_int
        fn.errcheck = _errcheck
    _lib, _lib_path = handle, path
    return handle


def lib():
    if _lib is None:
        load_library()
    return _lib


def library_path():
    return _lib_path


_initialised = False


def init(device_ordinal=None):
    """Select the GPU of this process (one process per GPU; LOCAL_RANK picks the ordinal)."""
    global _initialised
    if _initialised:
        return
    if device_ordinal is None:
        device_ordinal = int(os.environ.get("NANOGRAD_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    lib().ng_init(int(device_ordinal))
    _initialised = True


def is_available():
    """True when the library is built and a CUDA device answers."""
    try:
        init()
        return True
    except (DeviceLibraryError, OSError):
        return False


# ------------------------------------------------------------------ thin helpers -----------
def malloc(nbytes):
    if not _initialised:
        init()
    out = u64(0)
    _lib.ng_malloc(_byref(out), nbytes if nbytes > 0 else 1)
    return out.value


def free(ptr):
    if ptr and _lib is not None:
        _lib.ng_free(ptr)


def sync():
    init()
    _lib.ng_sync()


def launch_count():
    out = u64(0)
    lib().ng_launch_count(ctypes.byref(out))
    return out.value


def mem_stats():
    a, b, c = u64(0), u64(0), u64(0)
    lib().ng_mem_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return {"live_bytes": a.value, "pooled_bytes": b.value, "device_allocs": c.value}


def graph_capturing():
    flag = c_int(0)
    lib().ng_graph_capturing(ctypes.byref(flag))
    return bool(flag.value)


def device_info():
    init()
    sm, mj, mn, mem = c_int(0), c_int(0), c_int(0), u64(0)
    _lib.ng_device_info(ctypes.byref(sm), ctypes.byref(mj), ctypes.byref(mn), ctypes.byref(mem))
    return {"sm_count": sm.value, "cc": (mj.value, mn.value), "total_mem": mem.value}


def i64_array(values):
    return (c_int64 * len(values))(*[int(v) for v in values])


def i32_array(values):
    return (c_int32 * len(values))(*[int(v) for v in values])


def u64_array(values):
    return (c_uint64 * len(values))(*[int(v) for v in values])


class Event:
    """CUDA event on the compute stream (timing)."""

    def __init__(self):
        init()
        h = u64(0)
        _lib.ng_event_create(ctypes.byref(h))
        self.handle = h.value

    def record(self):
        _lib.ng_event_record(self.handle)
        return self

    def synchronize(self):
        _lib.ng_event_sync(self.handle)

    def elapsed_ms(self, stop):
        ms = c_float(0)
        _lib.ng_event_elapsed_ms(self.handle, stop.handle, ctypes.byref(ms))
        return ms.value

    def __del__(self):
        try:
            if _lib is not None and self.handle:
                _lib.ng_event_destroy(self.handle)
        except Exception:
            pass


def pinned_empty(shape, dtype="float32"):
    """NumPy array backed by page-locked host memory (async H2D source / D2H target)."""
    import numpy as np
    init()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if not isinstance(shape, int) else int(shape)
    nbytes = max(n * dt.itemsize, 1)
    p = c_void_p(0)
    _lib.ng_pinned_alloc(ctypes.byref(p), nbytes)
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = (p.value, buf)
    return arr


_PINNED = {}
