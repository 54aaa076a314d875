"""NumPy restatement of nanograd's CPU operators (forward and backward), function style.

Every function names the reference lines it restates.  The arithmetic intentionally reproduces
the reference's dtype behaviour — e.g. Add's backward multiplies by a float64 ``np.ones``
(ops_cpu.py:174-175), OneHot returns float64 (ops_cpu.py:43), Max's backward divides by an int64
count (ops_cpu.py:126-127) — because that is what the device path is compared against.

One deliberate difference: ``conv2d_forward`` uses the batch stride C*H*W.  The reference writes
C*H*H (ops_cpu.py:356), which is the same number for the square inputs every reference test and
every BASELINE config uses, and reads out of bounds otherwise; parity is defined on H == W.
"""
import numpy as np
from numpy.lib.stride_tricks import as_strided


# ------------------------------------------------------------------ helpers
def unbroadcast(grad, shape, to_keep=0):
    """ops_cpu.py:12-18 — sum leading axes away, then sum (keepdims) every axis whose size differs."""
    while grad.ndim != len(shape):
        grad = grad.sum(axis=0)
    for ax in range(len(shape) - to_keep):
        if grad.shape[ax] != shape[ax]:
            grad = grad.sum(axis=ax, keepdims=True)
    return grad


def pad_slice(a, indices):
    """ops_cpu.py:20-33 (inner_slice) — slice [start, stop) per axis; the part of the window that
    lies outside ``a`` reads as zero (negative start / stop past the end = padding)."""
    before_after = [(max(0, -lo), max(0, hi - a.shape[ax])) for ax, (lo, hi) in enumerate(indices)]
    padded = np.pad(a, before_after, mode="constant")
    window = tuple(slice(lo + before_after[ax][0], hi + before_after[ax][0]) for ax, (lo, hi) in enumerate(indices))
    return padded[window]


def conv_out_len(length, kernel, stride, padding=0):
    """conv_ops.py:4-42"""
    return int((length + 2 * padding - kernel) // stride + 1)


def _as_axis_list(axis):
    return [axis] if isinstance(axis, int) else axis


def _keep_shape(x, axis):
    return [1 if axis is None or i in axis else x.shape[i] for i in range(x.ndim)]


# ------------------------------------------------------------------ shape ops
def onehot_forward(labels, num_classes):
    """ops_cpu.py:41-45 — float64 result, integer-truncated labels."""
    idx = labels.astype(int)
    out = np.zeros((idx.shape[0], num_classes))
    out[np.arange(len(out)), idx] = 1
    return out


def slice_forward(x, indices):
    """ops_cpu.py:78-80"""
    return pad_slice(x, indices)


def slice_backward(grad, in_shape, indices):
    """ops_cpu.py:83-86 — the inverse window: crop what was padded, pad what was cropped."""
    inverse = [(0 - lo, grad.shape[ax] + (in_shape[ax] - hi)) for ax, (lo, hi) in enumerate(indices)]
    return pad_slice(grad, inverse)


# ------------------------------------------------------------------ reductions
def extreme_forward(x, axis, kind):
    """ops_cpu.py:113-119 (Max) / 132-138 (Min).  Returns (out, out_keepdims, axis_list)."""
    axis = _as_axis_list(axis)
    fn = np.amax if kind == "max" else np.amin
    keep = fn(x, axis=None if axis is None else tuple(axis), keepdims=True)
    out = keep
    if axis is not None:
        out = keep.reshape([x.shape[i] for i in range(x.ndim) if i not in axis])
    return out, keep, axis


def extreme_backward(grad, x, axis, keep):
    """ops_cpu.py:122-127 / 141-146 — equality mask, gradient shared equally among ties."""
    shape = _keep_shape(x, axis)
    mask = (x == keep.reshape(shape))
    count = mask.sum(axis=None if axis is None else tuple(axis), keepdims=True)
    return mask * grad.reshape(shape) / count


def sum_forward(x, axis=None):
    """ops_cpu.py:151-155 — axis=None gives shape (1,)."""
    if axis is None:
        return np.array([x.sum()])
    return x.sum(axis=axis)


def sum_backward(grad, x, axis):
    """ops_cpu.py:158-162"""
    axis = [axis] if type(axis) == int else axis
    return grad.reshape(_keep_shape(x, axis)) + np.zeros_like(x)


# ------------------------------------------------------------------ elementwise
def add_backward(grad, a_shape, b_shape):
    """ops_cpu.py:172-176 — note the float64 ``np.ones``."""
    return unbroadcast(grad * np.ones(a_shape), a_shape), unbroadcast(grad * np.ones(b_shape), b_shape)


def mul_backward(grad, a, b):
    """ops_cpu.py:186-190"""
    return unbroadcast(grad * b, a.shape), unbroadcast(grad * a, b.shape)


def matmul_backward(grad, a, b):
    """ops_cpu.py:200-204"""
    return grad @ b.T, a.T @ grad


def pow_backward(grad, x, power):
    """ops_cpu.py:248-251"""
    with np.errstate(all="ignore"):
        gx = unbroadcast(power * (x ** (power - 1.0)) * grad, x.shape)
        gp = unbroadcast((x ** power) * np.log(x) * grad, power.shape)
    return gx, gp


def sigmoid(x):
    """ops_cpu.py:9-10"""
    return 1 / (1 + np.exp(-x))


def relu_backward(grad, x):
    """ops_cpu.py:261-263 — (x >= 0): the subgradient at 0 is 1."""
    return (x >= 0) * grad


def sigmoid_backward(grad, x):
    """ops_cpu.py:273-275"""
    return grad * sigmoid(x) * (1 - sigmoid(x))


def tanh_backward(grad, x):
    """ops_cpu.py:285-287"""
    return grad * (1 - np.power(np.tanh(x), 2))


# ------------------------------------------------------------------ convolutions
def conv1d_forward(x, w, stride):
    """ops_cpu.py:292-318 — strided (C, S, N, OL) view of x, tensordot with w over (C, S).
    Returns (y, cols) where cols is the (N, C, OL, S) view the backward needs."""
    n, c, length = x.shape
    k, _, s = w.shape
    ol = conv_out_len(length, s, stride)
    item = x.itemsize
    view = as_strided(x, shape=(c, s, n, ol), strides=(length * item, item, c * length * item, stride * item),
                      writeable=False)
    cols = view.transpose(2, 0, 1, 3)                      # (N, C, S, OL)
    y = np.tensordot(cols, w.transpose(1, 2, 0), axes=[(1, 2), (0, 1)])   # (N, OL, K)
    return y.transpose(0, 2, 1), cols.transpose(0, 1, 3, 2)


def conv1d_backward(grad, x, w, stride, cols):
    """ops_cpu.py:321-342 — wgrad by tensordot over (N, OL); dgrad by a loop over output
    positions that scatter-adds grad[:, :, X] @ w."""
    n, c, length = x.shape
    _, _, s = w.shape
    ol = grad.shape[2]
    grad_w = np.tensordot(grad, cols, axes=[(0, 2), (0, 2)])
    grad_x = np.zeros((n, c, length), dtype=grad.dtype)
    for pos in range(ol):
        lo = pos * stride
        grad_x[:, :, lo:lo + s] += np.tensordot(grad[:, :, pos], w, axes=[(1), (0)])
    return grad_x, grad_w


def conv2d_forward(x, w, stride):
    """ops_cpu.py:347-375 — strided (C, R, S, N, OH, OW) view of x, tensordot with w over (C, R, S).
    Returns (y, cols) where cols is the (N, C, OH, OW, R, S) view the backward needs."""
    n, c, h, wd = x.shape
    k, _, r, s = w.shape
    oh, ow = conv_out_len(h, r, stride), conv_out_len(wd, s, stride)
    item = x.itemsize
    view = as_strided(
        x, shape=(c, r, s, n, oh, ow),
        strides=(h * wd * item, wd * item, item, c * h * wd * item, stride * wd * item, stride * item),
        writeable=False)
    cols = view.transpose(3, 0, 1, 2, 4, 5)                # (N, C, R, S, OH, OW)
    y = np.tensordot(cols, w.transpose(1, 2, 3, 0), axes=[(1, 2, 3), (0, 1, 2)])   # (N, OH, OW, K)
    return y.transpose(0, 3, 1, 2), cols.transpose(0, 1, 4, 5, 2, 3)


def conv2d_backward(grad, x, w, stride, cols):
    """ops_cpu.py:378-400 — wgrad by tensordot over (N, OH, OW); dgrad by a Python loop over
    all OH*OW output pixels, each a small GEMM scatter-added into the R x S input window."""
    n, c, h, wd = x.shape
    _, _, r, s = w.shape
    oh, ow = grad.shape[2], grad.shape[3]
    grad_w = np.tensordot(grad, cols, axes=[(0, 2, 3), (0, 2, 3)])
    grad_x = np.zeros((n, c, h, wd), dtype=grad.dtype)
    for pix in range(oh * ow):
        col, row = pix % ow, pix // ow
        left, top = col * stride, row * stride
        grad_x[:, :, top:top + r, left:left + s] += np.tensordot(grad[:, :, row, col], w, axes=[(1), (0)])
    return grad_x, grad_w
