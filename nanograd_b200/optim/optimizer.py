"""SGD / Adam / AdamW with the reference's constructor signatures, state attribute names and
update rules (nanograd/optim/optimizer.py:5-128).

On ``Device.GPU`` one multi-tensor kernel updates every parameter and its state **in place** per
``step()`` (the reference issues 5-15 dispatched elementwise ops per parameter and rebinds
``p.data`` to a fresh buffer; values are the same, but anything aliasing the parameter's buffer —
a Reshape view, a saved activation — sees the new values; ``Tensor.copy()`` snapshots).  When a
data-parallel group is active (nanograd_b200.dist) the gradients already live in one flat bucket
(``dist.GradBucket``) that is summed across ranks with NCCL while the backward pass is still
running, and the 1/world_size average is folded into the update kernel.

The step itself is written as functions over a duck-typed optimizer (``device_sgd_step`` ...)
so that ``nanograd_b200.integrate`` can bind the same code to the reference's own optimizer
classes.  This package ships the device step only: a step over host tensors is the reference's
job (the test-suite plugs the NumPy oracle in through ``set_host_step``).
"""
from nanograd_b200 import backend as B
from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


def _is_gpu(t):
    # compared by name: the reference's Device enum is a different object (nanograd/device.py)
    return t.device.name == "GPU"


def on_device(opt):
    return len(opt.params) > 0 and all(_is_gpu(p) for p in opt.params)


def _device_tables(opt, *state_lists):
    """Pointer tables for the multi-tensor kernels.  Parameters without a gradient are skipped
    (the reference would raise on ``lr * None``).  Gradients and optimizer state that still live
    on the host — the optimizer was built before ``model.gpu()``, the flow of the reference's
    tests/test_mnist.py:95-96 — are moved to the device here, once."""
    active = [i for i, p in enumerate(opt.params) if p.grad is not None]
    params = [opt.params[i] for i in active]
    for p in params:
        if not _is_gpu(p.grad):
            p.grad.gpu()
    for states in state_lists:
        for s in states:
            if not _is_gpu(s):
                s.gpu()
    numels = [p.data.size for p in params]
    grad_ptrs, grad_scale = _reduced_grad_ptrs(opt, params, numels)
    tables = [B.u64_array([p.data.ptr for p in params]), B.u64_array(grad_ptrs)]
    for states in state_lists:
        tables.append(B.u64_array([states[i].data.ptr for i in active]))
    return len(params), tables, B.i64_array(numels), grad_scale


def _mirror_scalar_shapes(opt):
    """A 0-d parameter (Swish's beta) comes out of the reference's update with shape (1,): the
    update multiplies by a Python scalar, which dispatch wraps in a 1-element tensor
    (tensor.py:482-483), and () broadcasts to (1,).  The in-place kernels keep the storage, so
    only the logical shape is adjusted."""
    for p in opt.params:
        if p.grad is not None and tuple(p.data.shape) == ():
            p.data.shape = (1,)


def _reduced_grad_ptrs(opt, params, numels):
    """Data parallel: the summed gradients are read from the flat bucket, scaled by 1/world."""
    from nanograd_b200 import dist
    if not dist.is_initialized() or dist.world_size() == 1:
        return [p.grad.data.ptr for p in params], 1.0
    bucket = getattr(opt, "_bucket", None)
    if bucket is None or not bucket.matches(params):
        bucket = opt._bucket = dist.GradBucket(params)
    return bucket.reduce(params), 1.0 / dist.world_size()


def device_sgd_step(opt):
    """m = m * momentum + lr * grad ; p = p - m  (optimizer.py:52-55), one launch."""
    n, (p, g, m), numels, scale = _device_tables(opt, opt.momentums)
    if n:
        B.lib().ng_multi_sgd(n, p, g, m, numels, float(opt.lr), float(opt.momentum), float(scale))
    _mirror_scalar_shapes(opt)


def device_adam_step(opt, decoupled_weight_decay=False):
    """Adam (optimizer.py:79-89) / AdamW (optimizer.py:115-128), one launch.  ``opt.t`` has
    already been advanced by the caller."""
    if B.graph_capturing():
        raise Exception("Adam/AdamW cannot be captured into a replayable step: the bias corrections 1 - beta^t are "
                        "host scalars that change every step (capture SGD steps, or run Adam eagerly)")
    n, (p, g, m, v), numels, scale = _device_tables(opt, opt.exp_avg, opt.exp_avg_sq)
    if n:
        wd = float(opt.weight_decay) if decoupled_weight_decay else 0.0
        B.lib().ng_multi_adam(n, p, g, m, v, numels, float(opt.lr), float(opt.beta1), float(opt.beta2),
                              float(opt.eps), float(1. - opt.beta1 ** opt.t), float(1. - opt.beta2 ** opt.t),
                              wd, 1 if decoupled_weight_decay else 0, float(scale))
    _mirror_scalar_shapes(opt)


_HOST_STEPS = {}


def set_host_step(kind, fn):
    """Plug a host (Device.CPU) implementation of ``kind`` in {"sgd", "adam", "adamw"} — used by
    the test-suite to run the same model code on the NumPy oracle (oracle/host_steps.py)."""
    _HOST_STEPS[kind] = fn


def _host_step(kind, opt):
    fn = _HOST_STEPS.get(kind)
    if fn is None:
        raise Exception("nanograd_b200 ships the Device.GPU optimizer step only; move the model with .gpu() "
                        "(or run the reference nanograd for CPU math).")
    fn(opt)


class Optimizer:
    """Base class (optimizer.py:5-34)."""

    def __init__(self, params: list):
        self.params = list(params)
        self.state = []
        self._bucket = None

    def step(self):
        raise NotImplementedError("Should be implemented in subclasses of Optimizer")

    def reset(self):
        raise NotImplementedError("Should be implemented in subclasses of Optimizer")

    def zero_grad(self):
        for param in self.params:
            param.grad = None

    def _zeros_like_params(self):
        return [Tensor.zeros(p.shape, device=self.params[0].device) for p in self.params]


class SGD(Optimizer):
    """optimizer.py:37-55"""

    def __init__(self, params: list, lr: float = 1e-3, momentum: float = 0) -> None:
        super(SGD, self).__init__(params)
        assert momentum >= 0.0, "Momentum can't be negative"
        self.lr, self.momentum = lr, momentum
        self.momentums = self._zeros_like_params()

    def step(self) -> None:
        if on_device(self):
            device_sgd_step(self)
        else:
            _host_step("sgd", self)


class Adam(Optimizer):
    """optimizer.py:58-89"""

    def __init__(self, params: list, lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999,
                 eps: float = 1e-8) -> None:
        super(Adam, self).__init__(params)
        assert (0 <= beta1) and (beta1 < 1), "Smoothing factor must be in [0,1)"
        assert (0 <= beta2) and (beta2 < 1), "Smoothing factor must be in [0,1)"
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.exp_avg = self._zeros_like_params()
        self.exp_avg_sq = self._zeros_like_params()

    def step(self) -> None:
        self.t += 1
        if on_device(self):
            device_adam_step(self)
        else:
            _host_step("adam", self)


class AdamW(Optimizer):
    """Adam with decoupled weight decay (optimizer.py:92-128)."""

    def __init__(self, params: list, lr: float = 1e-3, weight_decay: float = 1e-2, beta1: float = 0.9,
                 beta2: float = 0.999, eps: float = 1e-8) -> None:
        super(AdamW, self).__init__(params)
        assert (0 <= beta1) and (beta1 < 1), "Smoothing factor must be in [0,1)"
        assert (0 <= beta2) and (beta2 < 1), "Smoothing factor must be in [0,1)"
        self.lr, self.weight_decay, self.eps = lr, weight_decay, eps
        self.beta1, self.beta2 = beta1, beta2
        self.t = 0
        self.exp_avg = self._zeros_like_params()
        self.exp_avg_sq = self._zeros_like_params()

    def step(self):
        self.t += 1
        if on_device(self):
            device_adam_step(self, decoupled_weight_decay=True)
        else:
            _host_step("adamw", self)
