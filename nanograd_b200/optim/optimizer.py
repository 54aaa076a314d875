"""SGD / Adam / AdamW with the reference's constructor signatures, state attribute names and
update rules (nanograd/optim/optimizer.py:5-128).

On ``Device.GPU`` one multi-tensor kernel updates every parameter and its state in place per
``step()`` (the reference issues 5-15 dispatched elementwise ops per parameter and rebinds
``p.data`` to a fresh buffer; values are the same).  When a data-parallel group is active
(nanograd_b200.dist) the gradients are packed into one flat bucket, summed across ranks with a
single NCCL allreduce, and the 1/world_size average is folded into the update kernel.
On any other device the step is written in Tensor arithmetic exactly as the reference does.
"""
from nanograd_b200 import backend as B
from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


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

    # ------------------------------------------------------------------ device helpers
    def _on_device(self):
        return len(self.params) > 0 and all(p.device == Device.GPU for p in self.params)

    def _device_tables(self, *state_lists):
        """Pointer tables for the multi-tensor kernels.  Parameters without a gradient are
        skipped (the reference would raise on ``lr * None``)."""
        active = [i for i, p in enumerate(self.params) if p.grad is not None]
        params = [self.params[i] for i in active]
        for p in params:
            if p.grad.device != Device.GPU:
                p.grad.gpu()
        numels = [p.data.size for p in params]
        grad_ptrs, grad_scale = self._reduced_grad_ptrs(params, numels)
        tables = [B.u64_array([p.data.ptr for p in params]), B.u64_array(grad_ptrs)]
        for states in state_lists:
            tables.append(B.u64_array([states[i].data.ptr for i in active]))
        return len(params), tables, B.i64_array(numels), grad_scale

    def _mirror_scalar_shapes(self):
        """A 0-d parameter (Swish's beta) comes out of the reference's update with shape (1,): the
        update multiplies by a Python scalar, which dispatch wraps in a 1-element tensor
        (tensor.py:482-483), and () broadcasts to (1,).  The in-place kernels keep the storage, so
        only the logical shape is adjusted."""
        for p in self.params:
            if p.grad is not None and tuple(p.data.shape) == ():
                p.data.shape = (1,)

    def _reduced_grad_ptrs(self, params, numels):
        """Data parallel: pack -> one allreduce(sum) -> kernels read the bucket, scaled by 1/world."""
        from nanograd_b200 import dist
        if not dist.is_initialized() or dist.world_size() == 1:
            return [p.grad.data.ptr for p in params], 1.0
        from nanograd_b200.nn.buffer import DeviceBuffer
        total = sum(numels)
        if self._bucket is None or self._bucket.size != total:
            self._bucket = DeviceBuffer((total,))
        n = len(params)
        B.lib().ng_multi_pack(n, B.u64_array([p.grad.data.ptr for p in params]), B.i64_array(numels),
                              self._bucket.ptr)
        B.lib().ng_nccl_allreduce_sum(self._bucket.ptr, total)
        ptrs, off = [], 0
        for k in numels:
            ptrs.append(self._bucket.ptr + 4 * off)
            off += k
        return ptrs, 1.0 / dist.world_size()


class SGD(Optimizer):
    """m = m * momentum + lr * grad ; p = p - m  (optimizer.py:37-55)."""

    def __init__(self, params: list, lr: float = 1e-3, momentum: float = 0) -> None:
        super(SGD, self).__init__(params)
        assert momentum >= 0.0, "Momentum can't be negative"
        self.lr, self.momentum = lr, momentum
        self.momentums = [Tensor.zeros(p.shape, device=self.params[0].device) for p in self.params]

    def step(self) -> None:
        if self._on_device():
            n, (p, g, m), numels, scale = self._device_tables(self.momentums)
            if n:
                B.lib().ng_multi_sgd(n, p, g, m, numels, float(self.lr), float(self.momentum), float(scale))
            self._mirror_scalar_shapes()
            return
        for i, p in enumerate(self.params):
            self.momentums[i] = self.momentums[i] * self.momentum + self.lr * p.grad
            p.data = (p - self.momentums[i]).data


class Adam(Optimizer):
    """optimizer.py:58-89"""

    def __init__(self, params: list, lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999,
                 eps: float = 1e-8) -> None:
        super(Adam, self).__init__(params)
        assert (0 <= beta1) and (beta1 < 1), "Smoothing factor must be in [0,1)"
        assert (0 <= beta2) and (beta2 < 1), "Smoothing factor must be in [0,1)"
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.exp_avg = [Tensor.zeros(p.shape, device=self.params[0].device) for p in self.params]
        self.exp_avg_sq = [Tensor.zeros(p.shape, device=self.params[0].device) for p in self.params]

    def step(self) -> None:
        self.t += 1
        if self._on_device():
            n, (p, g, m, v), numels, scale = self._device_tables(self.exp_avg, self.exp_avg_sq)
            if n:
                B.lib().ng_multi_adam(n, p, g, m, v, numels, float(self.lr), float(self.beta1), float(self.beta2),
                                      float(self.eps), float(1. - self.beta1 ** self.t),
                                      float(1. - self.beta2 ** self.t), 0.0, 0, float(scale))
            self._mirror_scalar_shapes()
            return
        for i, p in enumerate(self.params):
            self.exp_avg[i] = self.beta1 * self.exp_avg[i] + (1. - self.beta1) * p.grad
            self.exp_avg_sq[i] = self.beta2 * self.exp_avg_sq[i] + (1. - self.beta2) * (p.grad ** 2)
            bias_correction1 = self.exp_avg[i] / (1. - (self.beta1 ** self.t))
            bias_correction2 = self.exp_avg_sq[i] / (1. - (self.beta2 ** self.t))
            p.data = (p - self.lr * bias_correction1 / (bias_correction2.sqrt() + self.eps)).data


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
        self.exp_avg = [Tensor.zeros(p.shape, device=self.params[0].device) for p in self.params]
        self.exp_avg_sq = [Tensor.zeros(p.shape, device=self.params[0].device) for p in self.params]

    def step(self):
        self.t += 1
        bias_correction1 = 1 - self.beta1 ** self.t
        bias_correction2 = 1 - self.beta2 ** self.t
        if self._on_device():
            n, (p, g, m, v), numels, scale = self._device_tables(self.exp_avg, self.exp_avg_sq)
            if n:
                B.lib().ng_multi_adam(n, p, g, m, v, numels, float(self.lr), float(self.beta1), float(self.beta2),
                                      float(self.eps), float(bias_correction1), float(bias_correction2),
                                      float(self.weight_decay), 1, float(scale))
            self._mirror_scalar_shapes()
            return
        for i, p in enumerate(self.params):
            self.exp_avg[i] = self.beta1 * self.exp_avg[i] + (1. - self.beta1) * p.grad
            self.exp_avg_sq[i] = self.beta2 * self.exp_avg_sq[i] + (1. - self.beta2) * (p.grad ** 2)
            denom = (self.exp_avg_sq[i] / bias_correction2).sqrt() + self.eps
            step_size = self.lr / bias_correction1
            p.data = (p * (1 - self.lr * self.weight_decay) - step_size * (self.exp_avg[i] / denom)).data
