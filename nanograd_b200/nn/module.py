"""Neural-network layers: the reference's ``nanograd/nn/module.py`` surface on the new backend.

Class names, constructor signatures, attribute names (``weight``, ``bias``, ``running_mean`` ...)
and numerical semantics follow the reference (module.py:10-863, cited per class).  Layers are
device-agnostic compositions of Tensor methods; where the tensor's device namespace offers a
fused op (``conv2dbias``, ``linear``, ``batchnorm2d``, ``nllloss`` — see nn/ops_cuda.py) the layer
issues that single op, otherwise it composes primitives exactly as the reference does
(BatchNorm excepted: device op only, see ``set_host_forward``).
"""
from typing import Union

import numpy as np

from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


def init_weights(shape: tuple, weight_initialization: str, fan_mode: str = "fan_in", **kwargs) -> Tensor:
    """Kaiming / Glorot initialisers (module.py:10-49).  As in the reference, ``in_features`` is
    ``shape[1]`` — for a conv filter that is C, not C*R*S — and the uniform variants keep
    NumPy's float64 until the tensor reaches the device."""
    assert type(shape) == tuple, f"Shape must be a tuple. Got {type(shape).__name__}."
    assert fan_mode in ["fan_in", "fan_out"], "Wrong fan mode. Only fan_in and fan_out supported."
    out_features, in_features = shape[0], shape[1]
    fan = in_features if fan_mode == "fan_in" else out_features

    if weight_initialization == "kaiming_normal":
        return Tensor.normal(0.0, np.sqrt(1.0 / fan), shape, **kwargs)
    if weight_initialization == "kaiming_uniform":
        bound = np.sqrt(3.0 / fan)
        return Tensor(np.random.uniform(-bound, bound, shape), **kwargs)
    if weight_initialization == "glorot_normal":
        return Tensor.normal(0.0, np.sqrt(2.0 / (in_features + out_features)), shape, **kwargs)
    if weight_initialization == "glorot_uniform":
        bound = np.sqrt(6.0 / (in_features + out_features))
        return Tensor(np.random.uniform(-bound, bound, shape), **kwargs)
    raise Exception("Unknown weight initialization methods. Only Glorot and Kaiming are available.")


def _fused(x: Tensor, name: str) -> bool:
    return name in Tensor.ops[x.device]


# BatchNorm has no host arithmetic in this package (the device op is one fused kernel; CPU math
# is the reference's job).  The test-suite plugs the reference's composition of primitives in
# here to run the same model code on the NumPy oracle (oracle/host_layers.py).
_HOST_FORWARD = {}


def set_host_forward(layer_name: str, fn) -> None:
    _HOST_FORWARD[layer_name] = fn


def _host_forward(layer, x: Tensor) -> Tensor:
    fn = _HOST_FORWARD.get(type(layer).__name__)
    if fn is None:
        raise Exception(f"nanograd_b200 ships the Device.GPU {type(layer).__name__} only; move the model and its "
                        "input with .gpu() (or run the reference nanograd for CPU math).")
    return fn(layer, x)


class Module:
    """Base class: tracks parameters, tensors and sub-modules through ``__setattr__``
    (module.py:52-158)."""

    def __init__(self) -> None:
        self._submodules = {}
        self._parameters = {}
        self._tensors = {}
        self.is_train = True

    def train(self):
        self.is_train = True
        return self

    def eval(self):
        self.is_train = False
        return self

    def forward(self, *args):
        raise NotImplementedError("Subclasses of module should have a forward method implemented")

    def is_parameter(self, obj) -> bool:
        return isinstance(obj, Tensor) and obj.is_parameter

    def parameters(self):
        self._ensure_is_initialized()
        for parameter in self._parameters.values():
            yield parameter
        for module in self._submodules.values():
            yield from module.parameters()

    def add_parameter(self, name, value) -> None:
        self._ensure_is_initialized()
        self._parameters[name] = value

    def add_module(self, name, value) -> None:
        self._ensure_is_initialized()
        self._submodules[name] = value

    def add_tensor(self, name, value) -> None:
        self._ensure_is_initialized()
        self._tensors[name] = value

    def __setattr__(self, name, value):
        if self.is_parameter(value):
            self.add_parameter(name, value)
            self.add_tensor(name, value)
        elif isinstance(value, Module):
            self.add_module(name, value)
        elif isinstance(value, Tensor):
            self.add_tensor(name, value)
        object.__setattr__(self, name, value)

    def __call__(self, *args):
        return self.forward(*args)

    def _ensure_is_initialized(self):
        for attr, what in (('_submodules', 'Modules'), ('_tensors', 'Tensors'), ('_parameters', 'Parameters')):
            if self.__dict__.get(attr) is None:
                raise Exception(f"{what} not initialized. Did not initialize the parent class. "
                                "Call super().__init__().")

    def _move(self, device: Device):
        for tensor in self._tensors.values():
            tensor.gpu() if device == Device.GPU else tensor.cpu()
        for module in self._submodules.values():
            module.gpu() if device == Device.GPU else module.cpu()
        return self

    def cpu(self):
        return self._move(Device.CPU)

    def gpu(self):
        return self._move(Device.GPU)


class Sequential(Module):
    """Runs layers in order (module.py:161-218).  As in the reference, ``train()``/``eval()``
    return None here while ``Module``'s return self."""

    def __init__(self, *layers) -> None:
        super().__init__()
        self.layers = layers
        for idx, layer in enumerate(self.layers):
            self.add_module(str(idx), layer)

    def __iter__(self):
        yield from self.layers

    def __getitem__(self, idx: int):
        return self.layers[idx]

    def train(self) -> None:
        self.is_train = True
        for submodule in self._submodules.values():
            submodule.train()

    def eval(self) -> None:
        self.is_train = False
        for submodule in self._submodules.values():
            submodule.eval()

    def forward(self, x: Tensor) -> Tensor:
        layers, i = self.layers, 0
        while i < len(layers):
            layer = layers[i]
            # BatchNorm2d directly followed by ReLU runs as one fused device op when the device
            # namespace provides it (same values as the two layers in sequence)
            if (type(layer) is BatchNorm2d and i + 1 < len(layers) and type(layers[i + 1]) is ReLU
                    and _fused(x, 'batchnorm2d')):
                x = layer.forward(x, relu=True)
                i += 2
                continue
            x = layer(x)
            i += 1
        return x

    def gpu(self):
        for layer in self.layers:
            layer.gpu()
        return self

    def cpu(self):
        for layer in self.layers:
            layer.cpu()
        return self


class Linear(Module):
    """y = x @ W.T + b with W stored (out_features, in_features) (module.py:220-269)."""

    def __init__(self, in_features: int, out_features: int, weight_initialization: str = "kaiming_normal",
                 fan_mode: str = "fan_in", with_bias=True) -> None:
        super().__init__()
        self.in_features, self.out_features, self.with_bias = in_features, out_features, with_bias
        self.weight = init_weights((self.out_features, self.in_features), weight_initialization, fan_mode,
                                   requires_grad=True, is_parameter=True, name="lin_weight")
        if self.with_bias:
            self.bias = Tensor.zeros((self.out_features,), requires_grad=True, is_parameter=True, name="lin_bias")

    def forward(self, x: Tensor) -> Tensor:
        if _fused(x, 'linear') and len(x.shape) == 2:
            out = x.linear(self.weight, self.bias) if self.with_bias else x.linear(self.weight)
        elif self.with_bias:
            out = x @ self.weight.T + self.bias.T
        else:
            out = x @ self.weight.T
        out.name = "lin_res"
        return out


class BatchNorm1d(Module):
    """Batch normalisation over (batch, features) (module.py:272-329)."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1) -> None:
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = Tensor.ones((self.num_features,), requires_grad=True, is_parameter=True, name="bn_gamma")
        self.bias = Tensor.zeros((self.num_features,), requires_grad=True, is_parameter=True, name="bn_beta")
        self.running_mean = Tensor.zeros((self.num_features,), requires_grad=False, is_parameter=False,
                                         name="bn_running_mean")
        self.running_var = Tensor.ones((self.num_features,), requires_grad=False, is_parameter=False,
                                       name="bn_running_var")

    def forward(self, x: Tensor) -> Tensor:
        if _fused(x, 'batchnorm2d') and len(x.shape) == 2:
            # (B, F) is the 2-D kernel's (N, C, HW=1) case
            out = x.batchnorm2d(self.weight, self.bias, running_mean=self.running_mean.data,
                                running_var=self.running_var.data, eps=self.eps, momentum=self.momentum,
                                training=bool(self.is_train))
            out.name = "bn_1d_res"
            return out
        return _host_forward(self, x)


class BatchNorm2d(Module):
    """Batch normalisation over (N, C, H, W) per channel (module.py:332-391)."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1) -> None:
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = Tensor.ones(self.num_features, requires_grad=True, is_parameter=True, name="bn_gamma")
        self.bias = Tensor.zeros(self.num_features, requires_grad=True, is_parameter=True, name="bn_beta")
        self.running_mean = Tensor.zeros(self.num_features, requires_grad=False, is_parameter=False,
                                         name="bn_running_mean")
        self.running_var = Tensor.ones(self.num_features, requires_grad=False, is_parameter=False,
                                       name="bn_running_var")

    def forward(self, x: Tensor, relu: bool = False) -> Tensor:
        """``relu`` is set by Sequential when the next layer is a ReLU (fused device op only)."""
        if _fused(x, 'batchnorm2d'):
            out = x.batchnorm2d(self.weight, self.bias, running_mean=self.running_mean.data,
                                running_var=self.running_var.data, eps=self.eps, momentum=self.momentum,
                                training=bool(self.is_train), relu=bool(relu))
            out.name = 'bn_2d_res'
            return out
        assert not relu, "the fused BatchNorm2d+ReLU path exists on the device namespace only"
        return _host_forward(self, x)


class Conv1d(Module):
    """1-D convolution layer (module.py:394-434).  padding = int or (before, after)."""

    def __init__(self, in_channel: int, out_channel: int, kernel_size: int, stride: int = 1,
                 padding: Union[tuple, int] = (0, 0), weight_initialization: str = "kaiming_normal") -> None:
        super().__init__()
        assert isinstance(padding, (tuple, int)), 'Wrong padding type. Must be integer or tuple of integers'
        if isinstance(padding, int):
            padding = (padding, padding)
        assert len(padding) == 2, 'Wrong padding dimensions. Padding must be an integer of a 2-dimensional tuple'
        self.in_channel, self.out_channel = in_channel, out_channel
        self.stride, self.padding = stride, padding
        self.kernel_size = kernel_size
        self.weight_initialization = weight_initialization
        shape = (self.out_channel, self.in_channel, self.kernel_size)
        self.weight = init_weights(shape, self.weight_initialization, requires_grad=True, is_parameter=True,
                                   name="conv_weight_1d")
        self.bias = Tensor.zeros(out_channel, requires_grad=True, is_parameter=True, name="conv_bias_1d")

    def forward(self, x: Tensor) -> Tensor:
        if _fused(x, 'conv1dbias'):
            return x.conv1dbias(self.weight, self.bias, stride=self.stride, padding=tuple(self.padding))
        x_padded = x.pad1d(self.padding)
        return x_padded.conv1d(self.weight, self.stride) + self.bias.reshape(shape=[1, -1, 1])


class Conv2d(Module):
    """2-D convolution layer (module.py:437-476).  padding = int or (left, right, top, bottom)."""

    def __init__(self, in_channel: int, out_channel: int, kernel_size: tuple, stride: tuple = 1,
                 padding: tuple = (0, 0, 0, 0), weight_initialization: str = "kaiming_normal") -> None:
        super().__init__()
        assert isinstance(padding, (tuple, int)), 'Wrong padding type. Must be integer or tuple of integers'
        if isinstance(padding, int):
            padding = (padding, padding, padding, padding)
        assert len(padding) == 4, 'Wrong padding dimensions'
        self.in_channel, self.out_channel = in_channel, out_channel
        self.kernel_size = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
        self.stride, self.padding = stride, padding
        self.weight_initialization = weight_initialization
        shape = (self.out_channel, self.in_channel, *self.kernel_size)
        self.weight = init_weights(shape, self.weight_initialization, requires_grad=True, is_parameter=True,
                                   name="conv_weight_2d")
        self.bias = Tensor.zeros(self.out_channel, requires_grad=True, is_parameter=True, name="conv_bias_2d")

    def forward(self, x: Tensor) -> Tensor:
        if _fused(x, 'conv2dbias'):
            return x.conv2dbias(self.weight, self.bias, stride=self.stride, padding=tuple(self.padding))
        x_padded = x.pad2d(self.padding)
        return x_padded.conv2d(self.weight, self.stride) + self.bias.reshape(shape=[1, -1, 1, 1])


class MaxPool1d(Module):
    """module.py:479-504"""

    def __init__(self, pool_size: int) -> None:
        super().__init__()
        self.pool_size = pool_size

    def forward(self, x: Tensor) -> Tensor:
        out = x.max_pool1d(self.pool_size)
        out.name = 'mpool1d_res'
        return out


class AvgPool1d(Module):
    """module.py:507-532"""

    def __init__(self, pool_size: int) -> None:
        super().__init__()
        self.pool_size = pool_size

    def forward(self, x: Tensor) -> Tensor:
        out = x.avg_pool1d(self.pool_size)
        out.name = 'avgpool1d_res'
        return out


class MaxPool2d(Module):
    """module.py:535-559"""

    def __init__(self, pool_size: tuple) -> None:
        super().__init__()
        self.pool_size = pool_size if isinstance(pool_size, tuple) else (pool_size, pool_size)

    def forward(self, x: Tensor) -> Tensor:
        out = x.max_pool2d(self.pool_size)
        out.name = 'mpool2d_res'
        return out


class AvgPool2d(Module):
    """module.py:562-586"""

    def __init__(self, pool_size: tuple) -> None:
        super().__init__()
        self.pool_size = pool_size if isinstance(pool_size, tuple) else (pool_size, pool_size)

    def forward(self, x: Tensor) -> Tensor:
        out = x.avg_pool2d(self.pool_size)
        out.name = 'avgpool2d_res'
        return out


class Flatten(Module):
    """(B, d1, d2, ...) -> (B, d1*d2*...) (module.py:589-624)."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        out = x.flatten()
        out.name = 'flatten_res'
        return out


class Dropout(Module):
    """module.py:627-659.  As in the reference the Bernoulli mask is sampled once, on the first
    training call, and reused afterwards; evaluation multiplies by (1 - p)."""

    def __init__(self, p: float = 0.5) -> None:
        super().__init__()
        self.mask, self.p = None, p

    def forward(self, x: Tensor) -> Tensor:
        if self.is_train:
            if self.mask is None:
                val = np.random.binomial(1, 1.0 - self.p, size=x.shape)
                self.mask = Tensor(val, name="dropout_mask", is_parameter=True, device=x.device)
            out = x * self.mask
            out.name = 'dropout_res'
            return out
        return x * (1 - self.p)


class NLLLoss(Module):
    """-(log_probs * onehot(target)).mean(): the mean runs over batch*classes, i.e. 1/classes of
    the usual NLL (module.py:662-686)."""

    def __init__(self) -> None:
        pass

    def forward(self, log_probs: Tensor, target: Tensor) -> Tensor:
        batch_size, num_classes = log_probs.shape
        if _fused(log_probs, 'nllloss'):
            return log_probs.nllloss(target)
        labels = target.onehot(num_classes=num_classes)
        return -(log_probs * labels).mean()

    __call__ = forward


class CrossEntropyLoss(Module):
    """log_softmax followed by NLLLoss — the loss the reference README names (README.md:45,98)
    but no longer ships at HEAD."""

    def __init__(self) -> None:
        pass

    def forward(self, logits: Tensor, target: Tensor) -> Tensor:
        return NLLLoss().forward(logits.log_softmax(), target)

    __call__ = forward


class MSELoss(Module):
    """((predicted - target) ** 2).mean() (module.py:689-712)."""

    def __init__(self) -> None:
        pass

    def forward(self, predicted: Tensor, target: Tensor) -> Tensor:
        if _fused(predicted, 'mseloss') and predicted.shape == target.shape:
            return predicted.mseloss(target)
        return ((predicted - target) ** 2).mean()

    __call__ = forward


# ---------------------------------------------------------------------------- activations
class ReLU(Module):
    """module.py:718-741"""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        return x.relu()


class LeakyReLU(Module):
    """relu(x) - alpha * relu(-x) (module.py:744-772)."""

    def __init__(self, alpha=0.01) -> None:
        super().__init__()
        self.alpha = Tensor(alpha, is_parameter=False, name='alpha_relu')
        self._alpha_value = float(alpha)

    def forward(self, x: Tensor) -> Tensor:
        if _fused(x, 'leakyrelu'):
            return x.leakyrelu(alpha=self._alpha_value)
        return x.relu() - self.alpha * (-x).relu()


class Sigmoid(Module):
    """module.py:775-794"""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        return x.sigmoid()


class Tanh(Module):
    """module.py:797-816"""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        return x.tanh()


class Swish(Module):
    """x * sigmoid(beta * x) with a learnt beta (module.py:818-841)."""

    def __init__(self, beta: float = 1.0) -> None:
        super().__init__()
        self.beta = Tensor(beta, requires_grad=True, is_parameter=True, name='swish_beta')

    def forward(self, x: Tensor) -> Tensor:
        if _fused(x, 'swish') and self.beta.device == x.device:
            return x.swish(self.beta)
        return x * (self.beta * x).sigmoid()


class Mish(Module):
    """x * tanh(log(1 + exp(x))) (module.py:844-864)."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        if _fused(x, 'mish'):
            return x.mish()
        return x * (1 + x.exp()).log().tanh()
