"""Tensor: data holder, autograd tape walk, composite ops and the per-device operator registry.

Same public surface as the reference's ``nanograd/tensor.py`` — constructor arguments,
``zeros/ones/arange/randn/normal/randint/eye``, ``to/cpu/gpu``, ``backward``, arithmetic dunders,
``mean/flatten/log_softmax``, pooling and padding helpers, ``plot_forward/plot_backward``,
``register/register_ops`` and the ``Tensor.ops[device][name]`` table (tensor.py:33-518) — so that
a model written against nanograd moves to the B200 by calling ``.gpu()`` and nothing else.

What differs underneath:
  * ``Device.GPU`` data is a ``DeviceBuffer`` in HBM served by the sm_100a library through
    ctypes (``backend``), not a PyOpenCL buffer; there is no CPU fallback on that path.
  * This package ships the ``Device.GPU`` namespace only.  ``Device.CPU`` tensors are plain NumPy
    holders (model construction, host<->device staging); CPU arithmetic is the reference's own
    job, and the test-suite registers the NumPy oracle there to compare the two devices through
    identical model code.
  * Composite methods (pooling, log_softmax) route to one fused device op when the tensor's
    device namespace provides it, and otherwise fall back to the reference's composition of
    primitives (tensor.py:327-331, 370-406).
  * Python scalars are passed as immediates (ScalarBuffer) instead of 4-byte device uploads.
  * ``build_graph_topology`` is not lru_cached (the reference's cache pins whole graphs — and
    on a GPU their activations — for up to 128 roots, tensor.py:210).
"""
from collections import defaultdict
from typing import Union

import numpy as np

from nanograd_b200.device import Device
from nanograd_b200.nn.buffer import DeviceBuffer, ScalarBuffer
from nanograd_b200.viz.comp_graph import BackwardGraphVisualizer, ForwardGraphVisualizer


class Tensor:
    ops = defaultdict(dict)  # ops[device][name] -> Function subclass

    def __init__(self,
                 data: Union[np.ndarray, DeviceBuffer, list, int, float],
                 requires_grad: bool = False,
                 is_parameter: bool = False,
                 device: Device = Device.CPU,
                 name: str = 'no_name',
                 op: str = None) -> None:
        self.data = self._move_data(data, device)
        self.requires_grad = requires_grad
        self.is_parameter = is_parameter
        self.device = device
        self.grad = None
        self.ctx = None
        self.name, self.op = name, op
        self.children = []

    # ------------------------------------------------------------------ properties
    @property
    def shape(self) -> tuple:
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def T(self):
        return self.transpose()

    # ------------------------------------------------------------------ initialisers
    @classmethod
    def zeros(cls, *shape, **kwargs):
        return cls(np.zeros(*shape, dtype=np.float32), **kwargs)

    @classmethod
    def ones(cls, *shape, **kwargs):
        return cls(np.ones(*shape, dtype=np.float32), **kwargs)

    @classmethod
    def arange(cls, *interval, **kwargs):
        return cls(np.arange(*interval).astype(np.float32), **kwargs)

    @classmethod
    def randn(cls, *shape, **kwargs):
        return cls(np.random.randn(*shape).astype(np.float32), **kwargs)

    @classmethod
    def normal(cls, loc, scale, *shape, **kwargs):
        return cls(np.random.normal(loc, scale, *shape).astype(np.float32), **kwargs)

    @classmethod
    def randint(cls, low, high, *shape, **kwargs):
        return cls(np.random.randint(low, high, *shape).astype(np.float32), **kwargs)

    @classmethod
    def eye(cls, dim, **kwargs):
        return cls(np.eye(dim).astype(np.float32), **kwargs)

    # ------------------------------------------------------------------ device movement
    @staticmethod
    def _move_data(data, device: Device):
        """Host<->device transfer hook (reference: tensor.py:153-182)."""
        assert device in Device, "Unsupported device. Only CPU and GPU available."

        if isinstance(data, DeviceBuffer):
            if device == Device.GPU:
                return data
            return data.numpy()  # blocking D2H

        if not isinstance(data, np.ndarray):
            data = np.array(data, dtype=np.float32)

        if device == Device.GPU:
            return DeviceBuffer(data.shape, hostbuf=data)  # H2D (cast to float32)
        return data

    def to(self, device: Device):
        self.data, self.device = self._move_data(self.data, device), device
        if self.grad is not None:
            self.grad.to(device)

    def cpu(self):
        self.to(Device.CPU)
        return self

    def gpu(self):
        from nanograd_b200 import backend
        backend.init()  # raises if the library is not built or no GPU answers
        if not Tensor.ops[Device.GPU]:
            raise Exception("No operators are registered for Device.GPU.")
        self.to(Device.GPU)
        return self

    def numpy(self) -> np.ndarray:
        """Host copy of the values (blocking for device tensors)."""
        return self.data.numpy() if isinstance(self.data, DeviceBuffer) else np.asarray(self.data)

    # ------------------------------------------------------------------ autograd
    def build_graph_topology(self):
        """Post-order list of the graph nodes that carry a ctx (iterative DFS)."""
        order, visited = [], set()
        stack = [(self, False)]
        while stack:
            node, expanded = stack.pop()
            if expanded:
                order.append(node)
                continue
            if id(node) in visited:
                continue
            visited.add(id(node))
            if node.ctx is not None:
                stack.append((node, True))
                for parent in reversed(node.ctx.parents):
                    if id(parent) not in visited:
                        stack.append((parent, False))
        return order

    def backward(self):
        if self.shape != (1,):
            raise Exception("Can't initiate backprop from a non scalar-valued tensor.")

        if isinstance(self.data, DeviceBuffer):
            # seed gradient written by a kernel (no host-to-device copy: the step stays capturable)
            from nanograd_b200 import backend
            seed = DeviceBuffer(self.shape)
            backend.lib().ng_fill(seed.ptr, 1.0, seed.size)
            self.grad = Tensor(seed, device=self.device, requires_grad=False)
        else:
            self.grad = Tensor.ones(self.shape, device=self.device, requires_grad=False)

        for node in reversed(self.build_graph_topology()):
            assert node.grad is not None, 'Got an unitialized gradient node'
            gradients = node.ctx.backward(node.ctx, node.grad.data)
            if len(node.ctx.parents) == 1 or not isinstance(gradients, (tuple, list)):
                gradients = [gradients]
            for tensor, grad in zip(node.ctx.parents, gradients):
                if grad is None:
                    continue
                assert grad.shape == tensor.shape, \
                    f"Mismatched tensor and grad shape. Got {grad.shape} and {tensor.shape}. " \
                    "Tensor and gradient should have the same shape."
                if tensor.grad is None:
                    tensor.grad = Tensor(grad, device=self.device, requires_grad=False)
                else:
                    tensor.grad = tensor.grad + Tensor(grad, device=self.device, requires_grad=False)

    def copy(self):
        """Independent snapshot of the values (tensor.py:245-246).  On Device.GPU the optimizers
        update parameters in place, so the copy owns its own allocation (one d2d copy)."""
        if isinstance(self.data, DeviceBuffer):
            from nanograd_b200 import backend
            dup = DeviceBuffer(self.data.shape)
            if dup.size:
                backend.lib().ng_d2d(dup.ptr, self.data.ptr, dup.size * 4)
            return Tensor(dup, device=self.device)
        return Tensor(np.array(self.data, copy=True), device=self.device)

    def __str__(self):
        return f"<NanoTensor({str(self.data)}, name={self.name}, device={self.device}>"

    __repr__ = __str__

    # ------------------------------------------------------------------ indexing
    def __getitem__(self, args):
        if args is None:
            args = []
        elif not isinstance(args, (list, tuple)):
            args = [args]
        if len(args) > len(self.shape):
            raise IndexError(f"too many indices for a tensor of shape {self.shape}")

        indices = []
        for i, arg in enumerate(args):
            if not isinstance(arg, slice):
                raise TypeError("Only slices are supported as indices")
            start, stop = arg.start, arg.stop
            if start is None:
                start = 0
            elif not isinstance(start, (int, np.integer)):
                raise TypeError(f"Indices must be integer. Got {type(start)}")
            if stop is None:
                stop = self.shape[i]
            elif not isinstance(stop, (int, np.integer)):
                raise TypeError(f"Indices must be integer. Got {type(stop)}")
            elif stop < 0:
                stop = self.shape[i] + stop
            assert arg.step is None or arg.step == 1, "Custom step not yet implemented"
            indices.append((int(start), int(stop)))
        # trailing axes are taken whole (the reference raises NameError here, tensor.py:285)
        indices += [(0, self.shape[i]) for i in range(len(args), len(self.shape))]
        return self.slice(indices=indices)

    # ------------------------------------------------------------------ derived arithmetic
    def __truediv__(self, other):
        return self * (other ** -1.0)

    def __rtruediv__(self, other):
        return other * (self ** -1.0)

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, other):
        return -self + other

    def sqrt(self):
        return self ** 0.5

    def mean(self, axis=None):
        out = self.sum(axis=axis)
        coeff = np.prod(out.shape) / np.prod(self.shape)
        return out * coeff

    def flatten(self):
        dim1, dim2 = self.shape[0], int(np.prod(self.shape[1:]))
        return self.reshape(shape=(dim1, dim2))

    # ------------------------------------------------------------------ activations / loss head
    def _has_op(self, name):
        return name in Tensor.ops[self.device]

    def log_softmax(self):
        if self._has_op('logsoftmax'):
            return self.logsoftmax()
        batch_size, num_classes = self.shape
        a = self.max(axis=1).reshape(shape=[batch_size, 1])  # log-sum-exp shift
        out = self - a - (self - a).exp().sum(axis=1).log().reshape(shape=[batch_size, 1])
        return out

    # ------------------------------------------------------------------ pooling / padding
    def _pool1d(self, field_length: int):
        x_unpadded = self[:, :, :self.shape[2] - self.shape[2] % field_length]
        shape = (x_unpadded.shape[0], x_unpadded.shape[1],
                 x_unpadded.shape[2] // field_length, field_length)
        return x_unpadded.reshape(shape=shape)

    def max_pool1d(self, kernel_size: int = 2):
        if self._has_op('maxpool2d') and len(self.shape) == 3:
            n, c, l = self.shape
            y = self.reshape(shape=(n, c, 1, l)).maxpool2d(pool_size=(1, kernel_size))
            return y.reshape(shape=(n, c, l // kernel_size))
        return self._pool1d(kernel_size).max(axis=3)

    def avg_pool1d(self, kernel_size: int = 2):
        if self._has_op('avgpool2d') and len(self.shape) == 3:
            n, c, l = self.shape
            y = self.reshape(shape=(n, c, 1, l)).avgpool2d(pool_size=(1, kernel_size))
            return y.reshape(shape=(n, c, l // kernel_size))
        return self._pool1d(kernel_size).mean(axis=3)

    def _pool2d(self, pool_size: tuple):
        x_unpadded = self[:, :,
                          :self.shape[2] - self.shape[2] % pool_size[0],
                          :self.shape[3] - self.shape[3] % pool_size[1]]
        shape = (x_unpadded.shape[0], x_unpadded.shape[1],
                 x_unpadded.shape[2] // pool_size[0], pool_size[0],
                 x_unpadded.shape[3] // pool_size[1], pool_size[1])
        return x_unpadded.reshape(shape=shape)

    def max_pool2d(self, pool_size: tuple = (2, 2)):
        if self._has_op('maxpool2d'):
            return self.maxpool2d(pool_size=tuple(pool_size))
        return self._pool2d(pool_size).max(axis=(3, 5))

    def avg_pool2d(self, pool_size: tuple = (2, 2)):
        if self._has_op('avgpool2d'):
            return self.avgpool2d(pool_size=tuple(pool_size))
        return self._pool2d(pool_size).mean(axis=(3, 5))

    def pad1d(self, pad: tuple):
        return self[:, :, -pad[0]:int(self.shape[2]) + pad[1]]

    def pad2d(self, pad: tuple):
        return self[:, :, -pad[2]:int(self.shape[2]) + pad[3], -pad[0]:int(self.shape[3]) + pad[1]]

    # ------------------------------------------------------------------ visualisation
    def plot_forward(self, rankdir="LR"):
        return ForwardGraphVisualizer().visualize(self, rankdir=rankdir)

    def plot_backward(self, rankdir="LR"):
        return BackwardGraphVisualizer().visualize(self, rankdir=rankdir)


# ---------------------------------------------------------------------- operator registry
def _wrap_scalar(value, like: Tensor) -> Tensor:
    if like.device == Device.GPU:
        return Tensor(ScalarBuffer(value), device=Device.GPU)
    return Tensor(np.array([value], dtype=like.dtype), device=like.device)


def register(name, operation, device=Device.CPU):
    """``Tensor.ops[device][name] = operation`` and a dispatching method/dunder on Tensor
    (reference: tensor.py:457-494)."""
    Tensor.ops[device][name] = operation

    def dispatch(*args, **kwargs):
        tensors = [arg for arg in args if isinstance(arg, Tensor)]
        if not tensors:
            raise TypeError(f"{name}: at least one argument must be a Tensor")
        first = tensors[0]
        args = [arg if isinstance(arg, Tensor) else _wrap_scalar(arg, first) for arg in args]
        ops = Tensor.ops[first.device]
        if name not in ops:
            if first.device == Device.CPU:
                raise Exception(
                    f"nanograd_b200 ships the Device.GPU operators only; '{name}' was called on a "
                    "Device.CPU tensor. Move the tensor with .gpu() (or run the reference nanograd for CPU math).")
            raise Exception(f"Operator '{name}' is not registered for {first.device}")
        op = ops[name]
        op.device = first.device
        return op.apply(op, *args, **kwargs)

    if name in ['add', 'mul', 'pow', 'matmul', 'neg']:
        setattr(Tensor, f"__{name}__", dispatch)
        if name != 'neg':
            setattr(Tensor, f"__r{name}__", lambda self, x: dispatch(x, self))
    else:
        setattr(Tensor, name, dispatch)


def register_ops(namespace, device=Device.CPU):
    """Register every Function subclass of ``namespace`` under its lower-cased class name
    (reference: tensor.py:497-500)."""
    import inspect
    from nanograd_b200.autograd import Function
    for cls_name, cls in inspect.getmembers(namespace, inspect.isclass):
        if cls_name[0] != "_" and issubclass(cls, Function) and cls is not Function:
            register(cls_name.lower(), cls, device=device)


def _register_device_namespace():
    from nanograd_b200.nn import ops_cuda
    register_ops(ops_cuda, device=Device.GPU)


_register_device_namespace()
