"""Device buffer held by ``Tensor.data`` on ``Device.GPU``.

Mirror of the reference's ``GPUBuffer`` (nanograd/nn/buffer.py:14-26): exposes ``.shape`` (tuple,
compared against gradients in Tensor.backward, tensor.py:238) and ``.dtype`` (always float32,
buffer.py:17); constructing one from another buffer aliases the same allocation under a new
shape (buffer.py:19-20 — how Reshape/Squeeze/Unsqueeze are free on the device).  The
allocation belongs to the base buffer and returns to the library's stream-ordered pool when
the last alias is garbage-collected.
"""
import ctypes
import math

import numpy as np

from nanograd_b200 import backend


def _numel(shape):
    return math.prod(map(int, shape))


class DeviceBuffer:
    def __init__(self, shape, hostbuf=None):
        # a training step builds ~110 of these: the common case (a tuple of Python ints, no host data) stays short
        if type(shape) is not tuple or not all(type(s) is int for s in shape):
            shape = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(map(int, shape))
        self.shape = shape
        self.dtype = np.float32
        self.size = size = math.prod(shape)
        self._base = None
        self._ptr = 0
        if hostbuf is None:
            self._ptr = backend.malloc(size * 4)
        elif isinstance(hostbuf, DeviceBuffer):
            if hostbuf.size != size:
                raise Exception(f"cannot view a buffer of {hostbuf.size} elements as shape {self.shape}")
            self._base = hostbuf._base if hostbuf._base is not None else hostbuf
            self._ptr = hostbuf.ptr
        else:
            self._ptr = backend.malloc(size * 4)
            host = np.ascontiguousarray(hostbuf, dtype=np.float32)
            if host.size != size:
                raise Exception(f"host array of {host.size} elements does not fill shape {self.shape}")
            if size:
                backend.lib().ng_h2d(self._ptr, host.ctypes.data_as(ctypes.c_void_p), size * 4)

    @property
    def ptr(self):
        return self._ptr

    @property
    def nbytes(self):
        return self.size * 4

    def view(self, shape):
        """Zero-copy alias with a new shape."""
        return DeviceBuffer(shape, hostbuf=self)

    @classmethod
    def alias_at(cls, base, offset_elems, shape):
        """Zero-copy alias of ``numel(shape)`` elements of ``base`` starting at ``offset_elems`` (slots of
        the data-parallel gradient arena)."""
        out = cls.__new__(cls)
        out.shape = tuple(map(int, shape))
        out.dtype = np.float32
        out.size = math.prod(out.shape)
        if offset_elems < 0 or offset_elems + out.size > base.size:
            raise Exception(f"alias of {out.size} elements at {offset_elems} does not fit a buffer of {base.size}")
        out._base = base._base if base._base is not None else base
        out._ptr = base.ptr + 4 * int(offset_elems)
        return out

    def numpy(self):
        """Blocking device-to-host copy (the only synchronising operation besides sync()).  A buffer that carries
        a ``ready`` event (recorded right after the kernel that produced it: the loss ops do) is read on a separate
        stream that waits for that event only, so ``float(loss.numpy())`` in a training loop does not drain the
        backward pass and the optimizer step already queued behind the loss."""
        out = np.empty(self.shape, dtype=np.float32)
        ready = getattr(self, "ready", None)
        if ready is not None:
            backend.lib().ng_d2h_after(out.ctypes.data_as(ctypes.c_void_p), self.ptr, self.size * 4, ready.handle)
        else:
            backend.lib().ng_d2h(out.ctypes.data_as(ctypes.c_void_p), self.ptr, self.size * 4)
        return out

    def mark_ready(self):
        """Record that everything this buffer will ever hold has been enqueued (see ``numpy``).  Not inside a
        stream capture: an event recorded there cannot be waited on from outside the graph."""
        if not backend.graph_capturing():
            self.ready = backend.Event().record()
        return self

    def copy_from_host(self, host):
        host = np.ascontiguousarray(host, dtype=np.float32)
        if host.size != self.size:
            raise Exception(f"host array of {host.size} elements does not fill shape {self.shape}")
        backend.lib().ng_h2d(self.ptr, host.ctypes.data_as(ctypes.c_void_p), self.size * 4)
        return self

    def __del__(self):
        try:
            if self._base is None and self._ptr:
                backend.free(self._ptr)
                self._ptr = 0
        except Exception:
            pass

    def __repr__(self):
        return f"<DeviceBuffer with shape {self.shape}>"


class LazyBuffer(DeviceBuffer):
    """A result whose device memory is only allocated and written if somebody asks for its pointer.

    BatchNorm2d's normalise pass (and the dx pass of its backward) produce tensors whose only reader, in a
    conv -> BatchNorm2d -> ReLU -> conv chain, is the staging copy of the next convolution's tensor-core
    operands.  The op returns one of these instead: ``source`` describes how to recompute the values
    (kind + the buffers involved), the convolution stages straight from it
    (``ng_conv2d_stage_f16_bn`` / ``_bnbwd``) and the tensor itself is never written.  Any other consumer
    (pooling, ``numpy()``, a view, an op of the generic namespace) reads ``.ptr``, which runs
    ``materialize(ptr)`` once — the kernel the eager path would have run — so values and semantics are those
    of a plain ``DeviceBuffer``."""

    def __init__(self, shape, materialize, kind=None, source=None):
        self.shape = tuple(map(int, shape))
        self.dtype = np.float32
        self.size = math.prod(self.shape)
        self._base = None
        self._ptr = 0
        self._materialize = materialize
        self.kind = kind
        self.source = source

    @property
    def pending(self):
        """True while the values exist only as a recipe."""
        return self._ptr == 0

    @property
    def ptr(self):
        if not self._ptr:
            self._ptr = backend.malloc(self.size * 4)
            fn, self._materialize = self._materialize, None
            fn(self._ptr)
        return self._ptr

    def materialize_with(self, fn):
        """Materialise through ``fn(ptr)`` instead of the default kernel (a consumer that wants a by-product of
        the same pass, e.g. channel sums).  No-op if the buffer was already written."""
        if self._ptr:
            return False
        self._ptr = backend.malloc(self.size * 4)
        self._materialize = None
        fn(self._ptr)
        return True


class ScalarBuffer(DeviceBuffer):
    """A Python scalar travelling as a 1-element tensor.

    The reference wraps every non-Tensor positional argument into ``Tensor(np.array([x]))`` on
    the input's device (tensor.py:482-483) — a 4-byte host-to-device copy per ``x * 0.5``.  Here
    the value stays on the host and rides in the kernel arguments; device memory is only
    materialised if some op really needs a pointer.  Once a pointer has been handed out the
    device copy is the truth (a kernel may have written through it): ``immediate`` turns False
    and ``numpy()`` reads the device.
    """

    def __init__(self, value):
        self.shape = (1,)
        self.dtype = np.float32
        self.size = 1
        self._base = None
        self._ptr = 0
        self.value = float(value)

    @property
    def immediate(self):
        """True while the value only lives on the host (usable as a kernel immediate)."""
        return self._ptr == 0

    @property
    def ptr(self):
        if not self._ptr:
            self._ptr = backend.malloc(4)
            host = np.array([self.value], dtype=np.float32)
            backend.lib().ng_h2d(self._ptr, host.ctypes.data_as(ctypes.c_void_p), 4)
        return self._ptr

    def numpy(self):
        if self._ptr:
            return DeviceBuffer.numpy(self)
        return np.array([self.value], dtype=np.float32)


class GPUBuffer(DeviceBuffer):
    """``DeviceBuffer`` under the reference's name and constructor signature
    ``GPUBuffer(ctx, shape, hostbuf=None)`` (nanograd/nn/buffer.py:14-26) — what
    ``nanograd/nn/buffer.py`` exports once the backend is bound.  ``ctx`` (the
    PyOpenCL context in the reference) is ignored."""

    def __init__(self, ctx, shape, hostbuf=None):
        DeviceBuffer.__init__(self, shape, hostbuf=hostbuf)
