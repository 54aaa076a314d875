"""The ``Device.GPU`` operator namespace: ``Function`` subclasses over ``DeviceBuffer``s.

This is the drop-in for the reference's PyOpenCL namespace ``nanograd/nn/ops_gpu.py``: the same
21 class names with the same forward signatures (GPU-side superset of defaults: ``stride=1``,
``axis=None``, ``axis=-1``, ``indices=None``; ops_gpu.py:389,405,561,600,623,646,666,749), registered
with ``register_ops(ops_cuda, device=Device.GPU)``.  Every forward/backward is one or a few calls
into the sm_100a library through ctypes (``backend``); nothing here computes on the host and
there is no CPU fallback.  Semantics follow the NumPy path ``nanograd/nn/ops_cpu.py`` (cited per
op), which is what the parity tests compare against.

On top of the 21 primitives the namespace provides fused ops that the device-agnostic layers
route to when present (Tensor.max_pool2d, Tensor.log_softmax, nn.Conv2d, nn.Linear, nn.NLLLoss,
nn.BatchNorm2d): ``conv2dbias``, ``conv1dbias``, ``linear``, ``maxpool2d``, ``avgpool2d``,
``logsoftmax``, ``nllloss``, ``batchnorm2d``.

Deviations from the reference that cannot change results for anything that requires a gradient:
backward skips the gradient of a parent whose ``requires_grad`` is False (the reference stores a
grad on every parent, tensor.py:236-243), and every device value is float32 (the NumPy path drifts
to float64 through Add.backward / OneHot / Max.backward, ops_cpu.py:43,126,174).
"""
import numpy as np

from nanograd_b200 import backend as B
from nanograd_b200 import dist as _dist
from nanograd_b200.autograd import Function
from nanograd_b200.nn.buffer import DeviceBuffer, LazyBuffer, ScalarBuffer

__all__ = []  # classes are discovered by register_ops via inspect


# ------------------------------------------------------------------------------------------
# helpers (underscore-prefixed: not registered)
# ------------------------------------------------------------------------------------------
def _lib():
    return B.lib()


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def _needs(ctx, i):
    """Does parent ``i`` need a gradient?  The reference computes every parent's gradient
    (tensor.py:236-243); here only tensors that ask for one — or are flagged is_parameter, like
    Dropout's mask (module.py:650-652), whose .grad the reference's users can read — get one."""
    if i >= len(ctx.parents):
        return False
    p = ctx.parents[i]
    return bool(p.requires_grad or getattr(p, "is_parameter", False))


def _param_out(ctx, i, shape):
    """Output buffer for the gradient of parent ``i``.  Under data parallelism a parameter's gradient
    is produced straight into its slot of the flat gradient arena (dist.GradBucket: no pack copy, and
    the slot's chunk can go to NCCL while backward continues); otherwise a fresh buffer.  Returns
    (buffer, token); call ``_dist.grad_slot_filled(token)`` once the kernel filling it is enqueued."""
    buf, token = _dist.grad_slot(ctx, i, shape)
    if buf is None:
        return DeviceBuffer(shape), None
    return buf, token


def _unary(op, a):
    out = DeviceBuffer(a.shape)
    _lib().ng_unary(op, out.ptr, a.ptr, a.size)
    return out


def _bcast_strides(shape, out_shape):
    """Element strides of a contiguous array of ``shape`` broadcast to ``out_shape``."""
    nd = len(out_shape)
    shape = (1,) * (nd - len(shape)) + tuple(shape)
    strides, st = [0] * nd, 1
    for i in range(nd - 1, -1, -1):
        if shape[i] == 1 and out_shape[i] != 1:
            strides[i] = 0
        else:
            if shape[i] != out_shape[i]:
                raise Exception(f"operands could not be broadcast together: {shape} vs {out_shape}")
            strides[i] = st if shape[i] != 1 else 0
        st *= shape[i]
    return strides


def _collapse(out_shape, sa, sb):
    """Merge neighbouring axes that are contiguous for both operands (fewer div/mods per element)."""
    dims = [(n, a, b) for n, a, b in zip(out_shape, sa, sb) if n != 1]
    if not dims:
        return [1], [0], [0]
    merged = [dims[0]]
    for n, a, b in dims[1:]:
        pn, pa, pb = merged[-1]
        if pa == a * n and pb == b * n:
            merged[-1] = (pn * n, a, b)
        else:
            merged.append((n, a, b))
    return [d[0] for d in merged], [d[1] for d in merged], [d[2] for d in merged]


def _binary(op, a, b):
    """out = f(a, b) with NumPy broadcasting; Python scalars ride as kernel immediates."""
    a_scalar = isinstance(a, ScalarBuffer) and a.immediate
    b_scalar = isinstance(b, ScalarBuffer) and b.immediate
    if b_scalar and not a_scalar:
        out = DeviceBuffer(a.shape if len(a.shape) >= 1 else (1,))
        _lib().ng_binary_scalar(op, out.ptr, a.ptr, b.value, 0, a.size)
        return out
    if a_scalar and not b_scalar:
        out = DeviceBuffer(b.shape if len(b.shape) >= 1 else (1,))
        _lib().ng_binary_scalar(op, out.ptr, b.ptr, a.value, 1, b.size)
        return out
    if a.shape == b.shape:
        out = DeviceBuffer(a.shape)
        n = B.i64_array([a.size])
        one = B.i64_array([1])
        _lib().ng_binary(op, out.ptr, a.ptr, b.ptr, 1, n, one, one)
        return out
    out_shape = tuple(int(s) for s in np.broadcast_shapes(a.shape, b.shape))
    sa, sb = _bcast_strides(a.shape, out_shape), _bcast_strides(b.shape, out_shape)
    cshape, csa, csb = _collapse(out_shape, sa, sb)
    if len(cshape) > 8:
        raise Exception("broadcast over more than 8 non-mergeable axes is not supported")
    out = DeviceBuffer(out_shape)
    _lib().ng_binary(op, out.ptr, a.ptr, b.ptr, len(cshape), B.i64_array(cshape), B.i64_array(csa), B.i64_array(csb))
    return out


def _norm_axes(axis, ndim):
    if axis is None:
        return list(range(ndim))
    if isinstance(axis, (int, np.integer)):
        axis = [int(axis)]
    return sorted({int(a) % ndim for a in axis})


def _reduce(op, x, axes, keepdims=False, out=None):
    """Reduce ``x`` over ``axes`` (list of non-negative ints)."""
    nd = len(x.shape)
    mask = [1 if i in axes else 0 for i in range(nd)]
    kept = [x.shape[i] for i in range(nd) if not mask[i]]
    full = [1 if mask[i] else x.shape[i] for i in range(nd)]
    if out is None:
        out = DeviceBuffer(kept if kept else (1,))
    if x.size == 0 and out.size:
        # an empty reduced axis: NumPy's sum is 0, its max / min have no identity and raise (ops_cpu.py:100-146)
        if op != B.R_SUM:
            raise ValueError("zero-size array to reduction operation %s which has no identity"
                             % ("maximum" if op == B.R_MAX else "minimum"))
        _lib().ng_memset_zero(out.ptr, out.size * 4)
    else:
        _lib().ng_reduce(op, out.ptr, x.ptr, nd, B.i64_array(x.shape), B.i32_array(mask))
    if keepdims:
        return out.view(full)
    return out


def _unbroadcast(grad, shape):
    """Sum ``grad`` down to ``shape`` (ops_cpu.py:12-18)."""
    shape = tuple(shape)
    if grad.shape == shape:
        return grad
    nd = len(grad.shape)
    padded = (1,) * (nd - len(shape)) + shape
    axes = [i for i in range(nd) if padded[i] != grad.shape[i]]
    lead = list(range(nd - len(shape)))
    axes = sorted(set(axes) | set(lead))
    out = _reduce(B.R_SUM, grad, axes, keepdims=True)
    return out.view(shape)


_ZERO = None


def _zero_scalar():
    global _ZERO
    if _ZERO is None:
        _ZERO = DeviceBuffer((1,), hostbuf=np.zeros(1, dtype=np.float32))
    return _ZERO


def _broadcast_to(buf, shape):
    """Materialise ``buf`` broadcast to ``shape``."""
    shape = tuple(shape)
    if buf.shape == shape:
        return buf
    if _numel(buf.shape) == _numel(shape):
        return buf.view(shape)
    # out = buf (broadcast strides) + 0 (stride 0 everywhere)
    sa = _bcast_strides(buf.shape, shape)
    cshape, csa, csb = _collapse(shape, sa, [0] * len(shape))
    out = DeviceBuffer(shape)
    _lib().ng_binary(B.B_ADD, out.ptr, buf.ptr, _zero_scalar().ptr, len(cshape), B.i64_array(cshape),
                     B.i64_array(csa), B.i64_array(csb))
    return out


def _scalar_int(v):
    """Positional non-Tensor arguments arrive as 1-element buffers (tensor.py:482-483)."""
    if isinstance(v, ScalarBuffer) and v.immediate:
        return int(v.value)
    if isinstance(v, DeviceBuffer):
        return int(v.numpy().reshape(-1)[0])
    if isinstance(v, np.ndarray):
        return int(v.reshape(-1)[0])
    return int(v)


# ------------------------------------------------------------------------------------------
# shape / indexing ops
# ------------------------------------------------------------------------------------------
class OneHot(Function):
    """ops_cpu.py:39-49 — labels (float-encoded ints) -> (rows, num_classes) one-hot."""

    @staticmethod
    def forward(ctx, input, num_classes):
        rows = input.size
        out = DeviceBuffer((rows, int(num_classes)))
        _lib().ng_onehot(out.ptr, input.ptr, rows, int(num_classes))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError


class Unsqueeze(Function):
    """ops_cpu.py:52-61"""

    @staticmethod
    def forward(ctx, input, axis=-1):
        ctx.save_for_backward(input.shape)
        shape = list(input.shape)
        ax = axis if axis >= 0 else len(shape) + 1 + axis
        shape.insert(ax, 1)
        return input.view(shape)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.view(ctx.saved_tensors[0])


class Squeeze(Function):
    """ops_cpu.py:64-73"""

    @staticmethod
    def forward(ctx, input, axis=-1):
        ctx.save_for_backward(input.shape)
        shape = list(input.shape)
        ax = axis % len(shape)
        if shape[ax] != 1:
            raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
        del shape[ax]
        return input.view(shape)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.view(ctx.saved_tensors[0])


def _slice(buf, indices):
    """inner_slice (ops_cpu.py:20-33): out-of-range parts read as zero (negative start = padding)."""
    starts = [int(p[0]) for p in indices]
    out_shape = [int(p[1]) - int(p[0]) for p in indices]
    if any(s < 0 for s in out_shape):
        raise Exception(f"invalid slice {indices} for shape {buf.shape}")
    out = DeviceBuffer(out_shape)
    if out.size:
        _lib().ng_slice(out.ptr, buf.ptr, len(out_shape), B.i64_array(buf.shape), B.i64_array(out_shape),
                        B.i64_array(starts))
    return out


class Slice(Function):
    """ops_cpu.py:76-86"""

    @staticmethod
    def forward(ctx, input, indices=None):
        ctx.save_for_backward(input.shape, indices)
        return _slice(input, indices)

    @staticmethod
    def backward(ctx, grad_output):
        shape, indices = ctx.saved_tensors
        inverse = [(0 - p[0], grad_output.shape[i] + (shape[i] - p[1])) for i, p in enumerate(indices)]
        return _slice(grad_output, inverse)


def _transpose(buf):
    nd = len(buf.shape)
    if nd < 2:
        return buf
    out = DeviceBuffer(tuple(reversed(buf.shape)))
    _lib().ng_permute(out.ptr, buf.ptr, nd, B.i64_array(buf.shape), B.i32_array(list(reversed(range(nd)))))
    return out


class Transpose(Function):
    """ops_cpu.py:89-96 — ``.T`` (all axes reversed)."""

    @staticmethod
    def forward(ctx, input):
        return _transpose(input)

    @staticmethod
    def backward(ctx, grad_output):
        return _transpose(grad_output)


class Reshape(Function):
    """ops_cpu.py:99-108 — zero-copy alias on the device (ops_gpu.py:587)."""

    @staticmethod
    def forward(ctx, input, shape):
        ctx.save_for_backward(input.shape)
        shape = [int(s) for s in shape]
        if -1 in shape:
            known = _numel([s for s in shape if s != -1])
            shape[shape.index(-1)] = input.size // max(known, 1)
        return input.view(shape)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.view(ctx.saved_tensors[0])


# ------------------------------------------------------------------------------------------
# reductions
# ------------------------------------------------------------------------------------------
def _extreme_forward(ctx, op, input, axis):
    axes = _norm_axes(axis, len(input.shape))
    out_keep = _reduce(op, input, axes, keepdims=True)
    ctx.save_for_backward(input, axes, out_keep)
    if axis is None:
        return out_keep  # np.amax(..., axis=None, keepdims=True): shape (1,)*ndim, ops_cpu.py:115-119
    return out_keep.view([input.shape[i] for i in range(len(input.shape)) if i not in axes])


def _extreme_backward(ctx, grad_output):
    # ops_cpu.py:121-127: mask = (input == out); grad = mask * g / mask.sum(axis)  → ties share
    input, axes, out_keep = ctx.saved_tensors
    mask = _binary(B.B_EQ, input, out_keep)
    count = _reduce(B.R_SUM, mask, axes, keepdims=True)
    share = _binary(B.B_DIV, grad_output.view(out_keep.shape), count)
    return _binary(B.B_MUL, mask, share)


class Max(Function):
    @staticmethod
    def forward(ctx, input, axis=None):
        return _extreme_forward(ctx, B.R_MAX, input, axis)

    @staticmethod
    def backward(ctx, grad_output):
        return _extreme_backward(ctx, grad_output)


class Min(Function):
    @staticmethod
    def forward(ctx, input, axis=None):
        return _extreme_forward(ctx, B.R_MIN, input, axis)

    @staticmethod
    def backward(ctx, grad_output):
        return _extreme_backward(ctx, grad_output)


class Sum(Function):
    """ops_cpu.py:149-162 — ``axis=None`` returns shape (1,)."""

    @staticmethod
    def forward(ctx, input, axis=None):
        axes = _norm_axes(axis, len(input.shape))
        ctx.save_for_backward(input.shape, axes)
        return _reduce(B.R_SUM, input, axes)

    @staticmethod
    def backward(ctx, grad_output):
        shape, axes = ctx.saved_tensors
        keep = [1 if i in axes else shape[i] for i in range(len(shape))]
        return _broadcast_to(grad_output.view(keep), shape)


# ------------------------------------------------------------------------------------------
# elementwise
# ------------------------------------------------------------------------------------------
class Add(Function):
    """ops_cpu.py:165-176"""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a.shape, b.shape)
        return _binary(B.B_ADD, a, b)

    @staticmethod
    def backward(ctx, grad_output):
        a_shape, b_shape = ctx.saved_tensors
        ga = _unbroadcast(grad_output, a_shape) if _needs(ctx, 0) else None
        gb = _unbroadcast(grad_output, b_shape) if _needs(ctx, 1) else None
        return ga, gb


class Mul(Function):
    """ops_cpu.py:179-190"""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _binary(B.B_MUL, a, b)

    @staticmethod
    def backward(ctx, grad_output):
        a, b = ctx.saved_tensors
        ga = _unbroadcast(_binary(B.B_MUL, grad_output, b), a.shape) if _needs(ctx, 0) else None
        gb = _unbroadcast(_binary(B.B_MUL, grad_output, a), b.shape) if _needs(ctx, 1) else None
        return ga, gb


class Pow(Function):
    """ops_cpu.py:241-251"""

    @staticmethod
    def forward(ctx, input, power):
        ctx.save_for_backward(input, power)
        return _binary(B.B_POW, input, power)

    @staticmethod
    def backward(ctx, grad_output):
        input, power = ctx.saved_tensors
        gi = gp = None
        if _needs(ctx, 0):
            gi = _unbroadcast(_binary(B.B_MUL, _binary(B.B_POW_DBASE, input, power), grad_output), input.shape)
        if _needs(ctx, 1):
            gp = _unbroadcast(_binary(B.B_MUL, _binary(B.B_POW_DEXP, input, power), grad_output), power.shape)
        return gi, gp


class Neg(Function):
    """ops_cpu.py:231-238"""

    @staticmethod
    def forward(ctx, input):
        return _unary(B.U_NEG, input)

    @staticmethod
    def backward(ctx, grad_output):
        return _unary(B.U_NEG, grad_output)


class Log(Function):
    """ops_cpu.py:207-216"""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return _unary(B.U_LOG, input)

    @staticmethod
    def backward(ctx, grad_output):
        return _binary(B.B_LOG_BWD, grad_output, ctx.saved_tensors[0])


class Exp(Function):
    """ops_cpu.py:219-228"""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return _unary(B.U_EXP, input)

    @staticmethod
    def backward(ctx, grad_output):
        return _binary(B.B_EXP_BWD, grad_output, ctx.saved_tensors[0])


class ReLU(Function):
    """ops_cpu.py:254-263 — backward is (input >= 0) * g: subgradient 1 at 0."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return _unary(B.U_RELU, input)

    @staticmethod
    def backward(ctx, grad_output):
        return _binary(B.B_RELU_BWD, grad_output, ctx.saved_tensors[0])


class Sigmoid(Function):
    """ops_cpu.py:266-275"""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return _unary(B.U_SIGMOID, input)

    @staticmethod
    def backward(ctx, grad_output):
        return _binary(B.B_SIGMOID_BWD, grad_output, ctx.saved_tensors[0])


class Tanh(Function):
    """ops_cpu.py:278-287"""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return _unary(B.U_TANH, input)

    @staticmethod
    def backward(ctx, grad_output):
        return _binary(B.B_TANH_BWD, grad_output, ctx.saved_tensors[0])


# ------------------------------------------------------------------------------------------
# contractions
# ------------------------------------------------------------------------------------------
def _gemm(a, b, trans_a=False, trans_b=False, bias=None, flags=0, out=None):
    """(batched) row-major op(a) @ op(b) (+ bias over the last axis)."""
    if len(a.shape) < 2 or len(b.shape) < 2:
        raise Exception(f"matmul needs operands with at least 2 axes, got {a.shape} and {b.shape}")
    batch_shape = a.shape[:-2]
    if len(b.shape) > 2 or len(a.shape) > 2:
        if a.shape[:-2] != b.shape[:-2]:
            raise Exception(f"batched matmul needs equal batch axes, got {a.shape} and {b.shape}")
    batch = _numel(batch_shape)
    am, ak = (a.shape[-1], a.shape[-2]) if trans_a else (a.shape[-2], a.shape[-1])
    bk, bn = (b.shape[-1], b.shape[-2]) if trans_b else (b.shape[-2], b.shape[-1])
    if ak != bk:
        raise Exception(f"matmul: inner dimensions differ ({a.shape} @ {b.shape}, trans_a={trans_a}, trans_b={trans_b})")
    if out is None:
        out = DeviceBuffer(tuple(batch_shape) + (am, bn))
    _lib().ng_gemm(out.ptr, a.ptr, b.ptr, am, bn, ak, int(trans_a), int(trans_b), max(batch, 1),
                   bias.ptr if bias is not None else 0, flags | B.math_flags())
    return out


class MatMul(Function):
    """ops_cpu.py:193-204 — backward uses transposed-operand GEMMs, no transpose kernels."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _gemm(a, b)

    @staticmethod
    def backward(ctx, grad_output):
        a, b = ctx.saved_tensors
        ga = gb = None
        if _needs(ctx, 0):
            out, tok = _param_out(ctx, 0, a.shape)
            ga = _gemm(grad_output, b, trans_b=True, out=out)
            _dist.grad_slot_filled(tok)
        if _needs(ctx, 1):
            out, tok = _param_out(ctx, 1, b.shape)
            gb = _gemm(a, grad_output, trans_a=True, out=out)
            _dist.grad_slot_filled(tok)
        return ga, gb


def _conv_geometry(x_shape, w_shape, stride, padding):
    """padding = (left, right, top, bottom), the order of Tensor.pad2d (tensor.py:419-429)."""
    n, c, h, w = x_shape
    k, cw, r, s = w_shape
    if c != cw:
        raise Exception(f"conv: input has {c} channels but the filter expects {cw}")
    pl, pr, pt, pb = padding
    oh = (h + pt + pb - r) // stride + 1  # conv_ops.py:39-42
    ow = (w + pl + pr - s) // stride + 1
    if oh <= 0 or ow <= 0:
        raise Exception(f"conv: filter {w_shape} does not fit input {x_shape} with padding {padding}")
    return n, c, h, w, k, r, s, oh, ow, pt, pl


_PLAN_CACHE = {}


def _tc_plan(geom, stride):
    """(fprop, dgrad, wgrad) -> does that pass run on the tensor-core kernels (current math mode)?"""
    key = (geom, stride, B.math_flags())
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        import ctypes
        n, c, h, wd, k, r, s, oh, ow, pt, pl = geom
        out = (ctypes.c_int * 4)()
        _lib().ng_conv2d_tc_plan(n, c, h, wd, k, r, s, stride, oh, ow, B.math_flags(), out)
        plan = _PLAN_CACHE[key] = (bool(out[0]), bool(out[1]), bool(out[2]), int(out[3]))
    return plan


def _stage_nhwc(buf, engine, channel_sums=False, sums_out=None):
    """Channels-last staging copy of an NCHW buffer for the tensor-core kernels (made once, shared
    by the passes of one layer that read the same tensor).  engine 1: FP32 NHWC; engine 2 (FP16x3):
    scaled (hi, lo) FP16 pairs plus 64 bytes of metadata.  channel_sums=True also returns the
    per-channel sums of `buf` from the same read (the bias gradient when `buf` is dy).

    Engine 2 uses what the producer left on the buffer: a pending ``LazyBuffer`` of a BatchNorm2d forward /
    backward is staged from its recipe (the tensor is never written), and a known magnitude (``absmax``: a
    one-word device buffer with the bits of max|buf| or an upper bound) replaces the absmax pass."""
    n, c = buf.shape[0], buf.shape[1]
    hw = _numel(buf.shape[2:])
    want_sums = channel_sums and n <= 65535
    if engine == 2:
        out = DeviceBuffer((n * c * hw + 16,))
        out.staged_flags = (B.math_flags(), engine)
        sums = (sums_out if sums_out is not None else DeviceBuffer((c,))) if want_sums else None
        sums_ptr = sums.ptr if sums is not None else 0
        kind = buf.kind if isinstance(buf, LazyBuffer) and buf.pending else None
        if kind == "bn_fwd" and not want_sums:
            x, gamma, beta, mean, invstd, act, absmax = buf.source
            _lib().ng_conv2d_stage_f16_bn(out.ptr, x.ptr, gamma.ptr, beta.ptr, mean.ptr, invstd.ptr, absmax.ptr,
                                          n, c, hw, act)
        elif kind == "bn_bwd":
            dy, x, gamma, beta, mean, invstd, bn_sums, act, bound, pooled = buf.source
            if pooled is not None and isinstance(dy, LazyBuffer) and dy.pending:
                dpool, pool_out, h, w = pooled
                _lib().ng_conv2d_stage_f16_bnbwd_pool(out.ptr, sums_ptr, dpool.ptr, pool_out.ptr, x.ptr, gamma.ptr, beta.ptr,
                                                      mean.ptr, invstd.ptr, bn_sums.ptr, bound.ptr, n, c, h, w, act)
            else:
                _lib().ng_conv2d_stage_f16_bnbwd(out.ptr, sums_ptr, dy.ptr, x.ptr, gamma.ptr, beta.ptr if act else 0,
                                                 mean.ptr, invstd.ptr, bn_sums.ptr, bound.ptr, n, c, hw, act)
        else:
            known = getattr(buf, "absmax", None)
            _lib().ng_conv2d_stage_f16(out.ptr, buf.ptr, sums_ptr, known.ptr if known is not None else 0, n, c, hw)
        if channel_sums:
            return out, sums
        return out
    out = DeviceBuffer(buf.shape)
    out.staged_flags = (B.math_flags(), engine)
    if want_sums:
        sums = sums_out if sums_out is not None else DeviceBuffer((c,))
        _lib().ng_conv2d_stage_nhwc_sum(out.ptr, buf.ptr, sums.ptr, n, c, hw, B.math_flags())
        return out, sums
    _lib().ng_conv2d_stage_nhwc(out.ptr, buf.ptr, n, c, hw, B.math_flags())
    return (out, None) if channel_sums else out


def _conv2d_forward(x, w, bias, stride, padding, flags=0, x_nhwc=None, w_staged=None):
    n, c, h, wd, k, r, s, oh, ow, pt, pl = _conv_geometry(x.shape, w.shape, stride, padding)
    y = DeviceBuffer((n, k, oh, ow))
    src, extra = (x_nhwc, B.F_X_NHWC) if x_nhwc is not None else (x, 0)
    wsrc, wextra = (w_staged, B.F_W_STAGED) if w_staged is not None else (w, 0)
    _lib().ng_conv2d_fprop(y.ptr, src.ptr, wsrc.ptr, bias.ptr if bias is not None else 0,
                           n, c, h, wd, k, r, s, stride, pt, pl, oh, ow, flags | B.math_flags() | extra | wextra)
    return y


def _conv2d_dgrad(dy, w, x_shape, stride, padding, dy_nhwc=None, w_staged=None):
    n, c, h, wd, k, r, s, oh, ow, pt, pl = _conv_geometry(x_shape, w.shape, stride, padding)
    dx = DeviceBuffer(x_shape)
    src, extra = (dy_nhwc, B.F_DY_NHWC) if dy_nhwc is not None else (dy, 0)
    wsrc, wextra = (w_staged, B.F_W_STAGED) if w_staged is not None else (w, 0)
    _lib().ng_conv2d_dgrad(dx.ptr, src.ptr, wsrc.ptr, n, c, h, wd, k, r, s, stride, pt, pl, oh, ow,
                           B.math_flags() | extra | wextra)
    return dx


def _stage_filters(w, want_dgrad):
    """FP16x3 engine: both staged layouts of the filters from one pass at the forward (the weights of a step serve
    its forward and its backward); returns (fprop copy, dgrad copy or None)."""
    k, c = w.shape[0], w.shape[1]
    rs = _numel(w.shape[2:])
    wf = DeviceBuffer((w.size + 16,))
    wd = DeviceBuffer((w.size + 16,)) if want_dgrad else None
    _lib().ng_conv2d_stage_filters_f16(wf.ptr, wd.ptr if wd is not None else 0, w.ptr, k, c, rs)
    return wf, wd


def _conv2d_wgrad(dy, x, w_shape, stride, padding, x_nhwc=None, dy_nhwc=None, out=None):
    n, c, h, wd, k, r, s, oh, ow, pt, pl = _conv_geometry(x.shape, w_shape, stride, padding)
    dw = out if out is not None else DeviceBuffer(w_shape)
    xs, fx = (x_nhwc, B.F_X_NHWC) if x_nhwc is not None else (x, 0)
    ds, fd = (dy_nhwc, B.F_DY_NHWC) if dy_nhwc is not None else (dy, 0)
    _lib().ng_conv2d_wgrad(dw.ptr, ds.ptr, xs.ptr, n, c, h, wd, k, r, s, stride, pt, pl, oh, ow,
                           B.math_flags() | fx | fd)
    return dw


def _conv_forward_shared(ctx, input, weight, bias, stride, padding):
    """Forward of Conv2d / Conv2dBias.  When both fprop and wgrad will run on tensor cores the NHWC
    staging copy of the input is made once here and kept for the backward."""
    geom = _conv_geometry(input.shape, weight.shape, stride, padding)
    f_tc, d_tc, w_tc, engine = _tc_plan(geom, stride)
    x_nhwc = None
    if f_tc and w_tc and _needs(ctx, 1):
        x_nhwc = _stage_nhwc(input, engine)
    ctx.conv_nhwc = x_nhwc
    w_f = None
    ctx.conv_w_dgrad = None
    if engine == 2 and f_tc:
        w_f, w_d = _stage_filters(weight, want_dgrad=d_tc and _needs(ctx, 0))
        if w_d is not None:
            w_d.staged_flags = (B.math_flags(), engine)
            ctx.conv_w_dgrad = w_d
    return _conv2d_forward(input, weight, bias, stride, padding, x_nhwc=x_nhwc, w_staged=w_f)


def _conv_backward_shared(ctx, grad_output, input, weight, stride, padding, want_db=False):
    geom = _conv_geometry(input.shape, weight.shape, stride, padding)
    f_tc, d_tc, w_tc, engine = _tc_plan(geom, stride)
    need_dx, need_dw = _needs(ctx, 0), _needs(ctx, 1)
    x_nhwc = getattr(ctx, "conv_nhwc", None)
    if x_nhwc is not None and (not w_tc or getattr(x_nhwc, "staged_flags", None) != (B.math_flags(), engine)):
        x_nhwc = None  # math mode changed between forward and backward: restage inside the call
    dy_nhwc = db = None
    db_out, db_tok = _param_out(ctx, 2, (weight.shape[0],)) if want_db else (None, None)
    if (need_dx and need_dw and d_tc and w_tc) or (engine == 2 and ((need_dx and d_tc) or (need_dw and w_tc))):
        # one staging pass of dy serves dgrad and wgrad, and the bias gradient rides on its read (FP16x3: staged
        # here even for a single pass, so that a BatchNorm recipe upstream is fused into the staging)
        dy_nhwc, db = _stage_nhwc(grad_output, engine, channel_sums=True, sums_out=db_out) if want_db \
            else (_stage_nhwc(grad_output, engine), None)
    if (want_db and db is None and dy_nhwc is None and isinstance(grad_output, LazyBuffer) and grad_output.pending
            and grad_output.kind == "bn_bwd"):
        # the gradient is a BatchNorm recipe and this convolution reads it as a plain NCHW tensor (SIMT / direct
        # kernels): materialise it with the pass that also yields its channel sums = this layer's bias gradient
        bdy, bx, gamma, beta, mean, invstd, bn_sums, act, _, _ = grad_output.source
        n_, c_ = grad_output.shape[0], grad_output.shape[1]
        hw_ = _numel(grad_output.shape[2:])
        if grad_output.materialize_with(lambda ptr: _lib().ng_bn2d_act_bwd_dx_sum(
                ptr, db_out.ptr, bdy.ptr, bx.ptr, gamma.ptr, beta.ptr if act else 0, mean.ptr, invstd.ptr, bn_sums.ptr,
                n_, c_, hw_, act)):
            db = db_out
    # wgrad first: it is what the data-parallel gradient arena is waiting for (dgrad only feeds this rank)
    dw = None
    if need_dw:
        dw_out, dw_tok = _param_out(ctx, 1, weight.shape)
        dw = _conv2d_wgrad(grad_output, input, weight.shape, stride, padding, x_nhwc=x_nhwc, dy_nhwc=dy_nhwc, out=dw_out)
    if want_db:
        if db is None:
            db = _channel_sum(grad_output, out=db_out)
        _dist.grad_slot_filled(db_tok)
    if need_dw:
        _dist.grad_slot_filled(dw_tok)
    w_d = getattr(ctx, "conv_w_dgrad", None)
    if w_d is not None and (not d_tc or w_d.staged_flags != (B.math_flags(), engine)):
        w_d = None  # math mode changed between forward and backward: restage inside the call
    dx = _conv2d_dgrad(grad_output, weight, input.shape, stride, padding, dy_nhwc=dy_nhwc, w_staged=w_d) if need_dx else None
    if want_db:
        return dx, dw, db
    return dx, dw


def _channel_sum(dy, out=None):
    """Bias gradient: sum over every axis but axis 1 (Add.backward's unbroadcast, ops_cpu.py:12-18)."""
    n, k = dy.shape[0], dy.shape[1]
    inner = _numel(dy.shape[2:])
    if out is None:
        out = DeviceBuffer((k,))
    _lib().ng_channel_sum(out.ptr, dy.ptr, n, k, inner)
    return out


class Conv2d(Function):
    """ops_cpu.py:345-400 — valid cross-correlation NCHW x KCRS, one stride for both axes."""

    @staticmethod
    def forward(ctx, input, weight, stride=1):
        stride = _scalar_int(stride)
        ctx.save_for_backward(input, weight, stride)
        return _conv_forward_shared(ctx, input, weight, None, stride, (0, 0, 0, 0))

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, stride = ctx.saved_tensors
        return _conv_backward_shared(ctx, grad_output, input, weight, stride, (0, 0, 0, 0))


def _as4d(buf):
    n, c, l = buf.shape
    return buf.view((n, c, 1, l))


class Conv1d(Function):
    """ops_cpu.py:290-342 — the 2-D kernels with H = R = 1."""

    @staticmethod
    def forward(ctx, input, weight, stride=1):
        stride = _scalar_int(stride)
        ctx.save_for_backward(input, weight, stride)
        y = _conv2d_forward(_as4d(input), _as4d(weight), None, stride, (0, 0, 0, 0))
        return y.view((y.shape[0], y.shape[1], y.shape[3]))

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, stride = ctx.saved_tensors
        dy = _as4d(grad_output)
        x4, w4 = _as4d(input), _as4d(weight)
        dx = dw = None
        if _needs(ctx, 0):
            dx = _conv2d_dgrad(dy, w4, x4.shape, stride, (0, 0, 0, 0)).view(input.shape)
        if _needs(ctx, 1):
            dw = _conv2d_wgrad(dy, x4, w4.shape, stride, (0, 0, 0, 0)).view(weight.shape)
        return dx, dw


# ------------------------------------------------------------------------------------------
# fused ops (new names; the device-agnostic layers use them when the namespace has them)
# ------------------------------------------------------------------------------------------
class Conv2dBias(Function):
    """nn.Conv2d.forward in one op: pad2d + conv2d + bias add (module.py:475-476).
    padding = (left, right, top, bottom)."""

    @staticmethod
    def forward(ctx, input, weight, bias, stride=1, padding=(0, 0, 0, 0)):
        stride = _scalar_int(stride)
        padding = tuple(int(p) for p in padding)
        ctx.save_for_backward(input, weight, stride, padding)
        return _conv_forward_shared(ctx, input, weight, bias, stride, padding)

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, stride, padding = ctx.saved_tensors
        if _needs(ctx, 2):
            return _conv_backward_shared(ctx, grad_output, input, weight, stride, padding, want_db=True)
        dx, dw = _conv_backward_shared(ctx, grad_output, input, weight, stride, padding)
        return dx, dw, None


class Conv1dBias(Function):
    """nn.Conv1d.forward in one op (module.py:433-434).  padding = (left, right)."""

    @staticmethod
    def forward(ctx, input, weight, bias, stride=1, padding=(0, 0)):
        stride = _scalar_int(stride)
        pad4 = (int(padding[0]), int(padding[1]), 0, 0)
        ctx.save_for_backward(input, weight, stride, pad4)
        y = _conv2d_forward(_as4d(input), _as4d(weight), bias, stride, pad4)
        return y.view((y.shape[0], y.shape[1], y.shape[3]))

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, stride, pad4 = ctx.saved_tensors
        dy = _as4d(grad_output)
        x4, w4 = _as4d(input), _as4d(weight)
        dx = dw = db = None
        if _needs(ctx, 0):
            dx = _conv2d_dgrad(dy, w4, x4.shape, stride, pad4).view(input.shape)
        if _needs(ctx, 1):
            dw = _conv2d_wgrad(dy, x4, w4.shape, stride, pad4).view(weight.shape)
        if _needs(ctx, 2):
            db = _channel_sum(dy)
        return dx, dw, db


class Linear(Function):
    """nn.Linear.forward in one op: x @ W.T (+ b) with W stored (out, in) (module.py:248-266).
    Called as x.linear(weight) or x.linear(weight, bias)."""

    @staticmethod
    def forward(ctx, input, weight, bias=None):
        ctx.save_for_backward(input, weight)
        return _gemm(input, weight, trans_b=True, bias=bias)

    @staticmethod
    def backward(ctx, grad_output):
        input, weight = ctx.saved_tensors
        dw = db = None
        if _needs(ctx, 1):
            out, tok = _param_out(ctx, 1, weight.shape)
            dw = _gemm(grad_output, input, trans_a=True, out=out)                  # (out,B)@(B,in)
            _dist.grad_slot_filled(tok)
        if len(ctx.parents) >= 3 and _needs(ctx, 2):
            out, tok = _param_out(ctx, 2, (grad_output.shape[1],))
            db = _reduce(B.R_SUM, grad_output, [0], out=out)
            _dist.grad_slot_filled(tok)
        dx = _gemm(grad_output, weight) if _needs(ctx, 0) else None               # (B,out)@(out,in)
        if len(ctx.parents) < 3:
            return dx, dw
        return dx, dw, db


def _pool_geometry(shape, pool_size):
    n, c, h, w = shape
    ph, pw = int(pool_size[0]), int(pool_size[1])
    if ph <= 0 or pw <= 0:
        raise Exception(f"pool size must be positive, got {pool_size}")
    return n, c, h, w, ph, pw


class MaxPool2d(Function):
    """Tensor.max_pool2d fused (tensor.py:370-398): crop, window max; the backward is
    Max.backward's tie-splitting equality mask (ops_cpu.py:121-127)."""

    @staticmethod
    def forward(ctx, input, pool_size=(2, 2)):
        n, c, h, w, ph, pw = _pool_geometry(input.shape, pool_size)
        out = DeviceBuffer((n, c, h // ph, w // pw))
        if isinstance(input, LazyBuffer) and input.pending and input.kind == "bn_fwd":
            # pool the BatchNorm (+ ReLU) output straight from its recipe: the activation is never written
            x, gamma, beta, mean, invstd, act, _ = input.source
            _lib().ng_maxpool2d_fwd_bn(out.ptr, x.ptr, gamma.ptr, beta.ptr, mean.ptr, invstd.ptr, n * c, c, h, w, ph, pw, act)
        else:
            _lib().ng_maxpool2d_fwd(out.ptr, input.ptr, n * c, h, w, ph, pw)
        if getattr(input, "absmax", None) is not None:
            out.absmax = input.absmax   # a window maximum never exceeds the tensor's largest magnitude
        ctx.save_for_backward(input, out, (ph, pw))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, out, (ph, pw) = ctx.saved_tensors
        n, c, h, w = input.shape
        if isinstance(input, LazyBuffer) and input.pending and input.kind == "bn_fwd":
            x, gamma, beta, mean, invstd, act, _ = input.source

            def pool_bwd(ptr, dy=grad_output):
                _lib().ng_maxpool2d_bwd_bn(ptr, dy.ptr, x.ptr, out.ptr, gamma.ptr, beta.ptr, mean.ptr, invstd.ptr,
                                           n * c, c, h, w, ph, pw, act)

            if (ph, pw) == (2, 2) and h % 2 == 0 and w % 8 == 0 and (h * w) % 32 == 0 and h % (32 // (16 if w % 16 == 0 else 8)) == 0:
                # the gradient of the (never written) BatchNorm output stays a recipe too: BatchNorm's backward
                # reduction and the staging pass of the convolution below recompute it per window
                return LazyBuffer(input.shape, pool_bwd, "pool_bwd", (grad_output, out, input))
            dx = DeviceBuffer(input.shape)
            pool_bwd(dx.ptr)
            return dx
        dx = DeviceBuffer(input.shape)
        _lib().ng_maxpool2d_bwd(dx.ptr, grad_output.ptr, input.ptr, out.ptr, n * c, h, w, ph, pw)
        return dx


class AvgPool2d(Function):
    """Tensor.avg_pool2d fused (tensor.py:400-406, mean = Sum * 1/(ph*pw), tensor.py:309-312)."""

    @staticmethod
    def forward(ctx, input, pool_size=(2, 2)):
        n, c, h, w, ph, pw = _pool_geometry(input.shape, pool_size)
        out = DeviceBuffer((n, c, h // ph, w // pw))
        _lib().ng_avgpool2d_fwd(out.ptr, input.ptr, n * c, h, w, ph, pw)
        if getattr(input, "absmax", None) is not None:
            out.absmax = input.absmax   # nor does a window mean
        ctx.save_for_backward(input.shape, (ph, pw))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        shape, (ph, pw) = ctx.saved_tensors
        n, c, h, w = shape
        dx = DeviceBuffer(shape)
        _lib().ng_avgpool2d_bwd(dx.ptr, grad_output.ptr, n * c, h, w, ph, pw)
        return dx


class LogSoftmax(Function):
    """Tensor.log_softmax fused (tensor.py:327-331): 2-D input, softmax over axis 1."""

    @staticmethod
    def forward(ctx, input):
        if len(input.shape) != 2:
            raise ValueError(f"log_softmax expects a 2-D tensor, got shape {input.shape}")
        rows, cols = input.shape
        out = DeviceBuffer(input.shape)
        _lib().ng_logsoftmax_fwd(out.ptr, input.ptr, rows, cols)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out = ctx.saved_tensors[0]
        rows, cols = out.shape
        dx = DeviceBuffer(out.shape)
        _lib().ng_logsoftmax_bwd(dx.ptr, grad_output.ptr, out.ptr, rows, cols)
        return dx


class NLLLoss(Function):
    """nn.NLLLoss.forward fused: -(log_probs * onehot(target)).mean() over rows*classes
    (module.py:684-686).  Result has shape (1,) like Sum(axis=None)."""

    @staticmethod
    def forward(ctx, log_probs, target):
        rows, cols = log_probs.shape
        if target.size != rows:
            raise Exception(f"NLLLoss: {rows} rows of log-probabilities but {target.size} targets")
        out = DeviceBuffer((1,))
        _lib().ng_nll_fwd(out.ptr, log_probs.ptr, target.ptr, rows, cols)
        ctx.save_for_backward(target, (rows, cols))
        return out.mark_ready()   # reading the loss waits for the forward pass only

    @staticmethod
    def backward(ctx, grad_output):
        target, (rows, cols) = ctx.saved_tensors
        if not _needs(ctx, 0):
            return None, None
        dlogp = DeviceBuffer((rows, cols))
        _lib().ng_nll_bwd(dlogp.ptr, grad_output.ptr, target.ptr, rows, cols)
        return dlogp, None


class MSELoss(Function):
    """nn.MSELoss.forward fused: ((predicted - target) ** 2).mean() (module.py:689-712).
    Result has shape (1,) like Sum(axis=None)."""

    @staticmethod
    def forward(ctx, predicted, target):
        if tuple(predicted.shape) != tuple(target.shape):
            raise Exception(f"MSELoss: shapes {predicted.shape} and {target.shape} differ")
        out = DeviceBuffer((1,))
        _lib().ng_mse_fwd(out.ptr, predicted.ptr, target.ptr, predicted.size)
        ctx.save_for_backward(predicted, target)
        return out.mark_ready()   # reading the loss waits for the forward pass only

    @staticmethod
    def backward(ctx, grad_output):
        predicted, target = ctx.saved_tensors
        grads = []
        for i in range(2):
            if not _needs(ctx, i):
                grads.append(None)
                continue
            d = DeviceBuffer(predicted.shape)
            _lib().ng_mse_bwd(d.ptr, grad_output.ptr, predicted.ptr, target.ptr, predicted.size, i)
            grads.append(d)
        return tuple(grads)


def _act_forward(kind, x, param_ptr=0, param=0.0):
    out = DeviceBuffer(x.shape)
    _lib().ng_act_fwd(kind, out.ptr, x.ptr, param_ptr, float(param), x.size)
    return out


def _act_backward(kind, grad, x, param_ptr=0, param=0.0):
    dx = DeviceBuffer(x.shape)
    _lib().ng_act_bwd(kind, dx.ptr, grad.ptr, x.ptr, param_ptr, float(param), x.size)
    return dx


class LeakyReLU(Function):
    """nn.LeakyReLU.forward fused: relu(x) - alpha * relu(-x) (module.py:744-772); the gradient at
    x = 0 is 1 + alpha, as the composition of two ReLU.backward's gives (ops_cpu.py:263)."""

    @staticmethod
    def forward(ctx, input, alpha=0.01):
        ctx.save_for_backward(input, float(alpha))
        return _act_forward(B.ACT_LEAKY_RELU, input, param=alpha)

    @staticmethod
    def backward(ctx, grad_output):
        input, alpha = ctx.saved_tensors
        return _act_backward(B.ACT_LEAKY_RELU, grad_output, input, param=alpha)


class Mish(Function):
    """nn.Mish.forward fused: x * tanh(log(1 + exp(x))) (module.py:844-864)."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return _act_forward(B.ACT_MISH, input)

    @staticmethod
    def backward(ctx, grad_output):
        return _act_backward(B.ACT_MISH, grad_output, ctx.saved_tensors[0])


class Swish(Function):
    """nn.Swish.forward fused: x * sigmoid(beta * x) with the learnt scalar beta read on the device
    (module.py:818-841); backward returns (dx, dbeta) with dbeta shaped like beta."""

    @staticmethod
    def forward(ctx, input, beta):
        if beta.size != 1:
            raise Exception(f"Swish: beta must hold one value, got shape {beta.shape}")
        ctx.save_for_backward(input, beta)
        return _act_forward(B.ACT_SWISH, input, param_ptr=beta.ptr)

    @staticmethod
    def backward(ctx, grad_output):
        input, beta = ctx.saved_tensors
        dx = _act_backward(B.ACT_SWISH, grad_output, input, param_ptr=beta.ptr) if _needs(ctx, 0) else None
        dbeta = None
        if _needs(ctx, 1):
            dbeta = DeviceBuffer(tuple(beta.shape))
            _lib().ng_swish_dbeta(dbeta.ptr, grad_output.ptr, input.ptr, beta.ptr, input.size)
        return dx, dbeta


def _bn_split_ok(shape):
    """Run BatchNorm2d as statistics now + element-wise pass on demand (LazyBuffer)?  Only when a consumer
    could fuse the element-wise pass: the FP16x3 convolution engine is enabled, the channel count is one it
    takes, the tensor is image-shaped, and BatchNorm is not synchronised across ranks (whose one-call kernels
    carry the collectives).  (Under the host emulation of the CPU test-suite no convolution fuses, so every
    LazyBuffer is materialised: the split kernels are exercised against the golden vectors there too.)"""
    return (len(shape) == 4 and shape[1] % 64 == 0 and (B.math_flags() & B.F_F16X3) != 0
            and not _dist.sync_bn())


class BatchNorm2d(Function):
    """nn.BatchNorm2d.forward fused (module.py:356-391).  ``running_mean`` / ``running_var`` are
    raw buffers updated in place in training mode; parents are (x, gamma, beta).  ``relu=True``
    also applies the ReLU that follows the layer (ops_cpu.py:254-263) in the same pass; the
    backward then masks the incoming gradient with (z >= 0) recomputed from x."""

    @staticmethod
    def forward(ctx, input, gamma, beta, running_mean=None, running_var=None, eps=1e-5, momentum=0.1,
                training=True, relu=False):
        n, c = input.shape[0], input.shape[1]
        hw = _numel(input.shape[2:])
        act = 1 if relu else 0
        if training:
            save_mean, save_invstd = DeviceBuffer((c,)), DeviceBuffer((c,))
            rm = running_mean.ptr if running_mean is not None else 0
            rv = running_var.ptr if running_var is not None else 0
            if _bn_split_ok(input.shape):
                # statistics now; the normalise pass only if somebody other than an FP16x3 convolution reads y
                absmax = DeviceBuffer((1,))
                # the recipe keeps its own copies of gamma / beta (written by the statistics kernel): the optimizers
                # update parameters in place, and an activation read after optimizer.step() must still be this step's
                snap = DeviceBuffer((2 * c,))
                _lib().ng_bn2d_act_fwd_stats(input.ptr, gamma.ptr, beta.ptr, rm, rv, save_mean.ptr, save_invstd.ptr,
                                             absmax.ptr, snap.ptr, n, c, hw, float(eps), float(momentum), act)
                gamma_s, beta_s = DeviceBuffer.alias_at(snap, 0, (c,)), DeviceBuffer.alias_at(snap, c, (c,))

                def apply(ptr, x=input):
                    _lib().ng_bn2d_act_apply(ptr, x.ptr, gamma_s.ptr, beta_s.ptr, save_mean.ptr, save_invstd.ptr, n, c, hw, act)

                out = LazyBuffer(input.shape, apply, "bn_fwd", (input, gamma_s, beta_s, save_mean, save_invstd, act, absmax))
                out.absmax = absmax
            else:
                out = DeviceBuffer(input.shape)
                _lib().ng_bn2d_act_fwd_train(out.ptr, input.ptr, gamma.ptr, beta.ptr, rm, rv,
                                             save_mean.ptr, save_invstd.ptr, n, c, hw, float(eps), float(momentum), act)
        else:
            out = DeviceBuffer(input.shape)
            _lib().ng_bn2d_act_fwd_eval(out.ptr, input.ptr, gamma.ptr, beta.ptr, running_mean.ptr, running_var.ptr,
                                        n, c, hw, float(eps), act)
            # evaluation mode treats the running statistics as constants
            save_mean = running_mean
            save_invstd = _unary(B.U_RECIP, _unary(B.U_SQRT, _binary(B.B_ADD, running_var, ScalarBuffer(eps))))
        ctx.save_for_backward(input, gamma, beta, save_mean, save_invstd, bool(training), act, out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, gamma, beta, save_mean, save_invstd, training, act, out = ctx.saved_tensors
        n, c = input.shape[0], input.shape[1]
        hw = _numel(input.shape[2:])
        if training:
            (dgamma, tok_g), (dbeta, tok_b) = _param_out(ctx, 1, (c,)), _param_out(ctx, 2, (c,))
            if _bn_split_ok(input.shape) and _needs(ctx, 0):
                # reductions now; the dx pass only if somebody other than an FP16x3 convolution reads dx
                bn_sums, bound = DeviceBuffer((2 * c + 1,)), DeviceBuffer((1,))
                if isinstance(out, LazyBuffer) and out.kind == "bn_fwd":
                    # the recipe below reads the forward's private copies of gamma / beta (same values now; the
                    # parameters themselves change in place at the next optimizer step)
                    gamma, beta = out.source[1], out.source[2]
                pooled = None
                if (isinstance(grad_output, LazyBuffer) and grad_output.pending and grad_output.kind == "pool_bwd"
                        and grad_output.source[2] is out):
                    # the incoming gradient is the recipe of a 2x2 max-pool backward over THIS layer's output: fused
                    dpool, pool_out, _ = grad_output.source
                    pooled = (dpool, pool_out, input.shape[2], input.shape[3])
                    _lib().ng_bn2d_act_bwd_sums_pool(dgamma.ptr, dbeta.ptr, bn_sums.ptr, bound.ptr, dpool.ptr, pool_out.ptr,
                                                     input.ptr, gamma.ptr, beta.ptr, save_mean.ptr, save_invstd.ptr,
                                                     n, c, input.shape[2], input.shape[3], act)
                else:
                    _lib().ng_bn2d_act_bwd_sums(dgamma.ptr, dbeta.ptr, bn_sums.ptr, bound.ptr, grad_output.ptr, input.ptr,
                                                gamma.ptr, beta.ptr, save_mean.ptr, save_invstd.ptr, n, c, hw, act)

                def dx_pass(ptr, dy=grad_output, x=input):
                    _lib().ng_bn2d_act_bwd_dx(ptr, dy.ptr, x.ptr, gamma.ptr, beta.ptr, save_mean.ptr, save_invstd.ptr,
                                              bn_sums.ptr, n, c, hw, act)

                dx = LazyBuffer(input.shape, dx_pass, "bn_bwd",
                                (grad_output, input, gamma, beta, save_mean, save_invstd, bn_sums, act, bound, pooled))
                dx.absmax = bound
            else:
                dx = DeviceBuffer(input.shape)
                _lib().ng_bn2d_act_bwd(dx.ptr, dgamma.ptr, dbeta.ptr, grad_output.ptr, input.ptr, gamma.ptr, beta.ptr,
                                       save_mean.ptr, save_invstd.ptr, n, c, hw, act)
            _dist.grad_slot_filled(tok_g)
            _dist.grad_slot_filled(tok_b)
            return dx, dgamma, dbeta
        # constants statistics: y = (x - m) * g * inv + b
        bshape = (1, c) + (1,) * (len(input.shape) - 2)
        scale = _binary(B.B_MUL, gamma, save_invstd).view(bshape)
        xhat = _binary(B.B_MUL, _binary(B.B_SUB, input, save_mean.view(bshape)), save_invstd.view(bshape))
        if act:
            # relu'(z) with z = xhat * gamma + beta (the reference's >= 0 subgradient)
            z = _binary(B.B_ADD, _binary(B.B_MUL, xhat, gamma.view(bshape)), beta.view(bshape))
            grad_output = _binary(B.B_RELU_BWD, grad_output, z)
        dx = _binary(B.B_MUL, grad_output, scale)
        axes = [0] + list(range(2, len(input.shape)))
        dgamma = _reduce(B.R_SUM, _binary(B.B_MUL, grad_output, xhat), axes)
        dbeta = _reduce(B.R_SUM, grad_output, axes)
        return dx, dgamma, dbeta
