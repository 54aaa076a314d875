"""TEST INFRASTRUCTURE — the reference's *compositions* of primitives, for host (Device.CPU) runs.

``nanograd_b200`` ships the Device.GPU side only: its BatchNorm layers and optimizers issue one
fused device op and have no host arithmetic of their own.  To run the same model code on the
NumPy oracle (tests, ``smoke()``, the CPU-baseline legs of bench.py) the reference's Tensor-level
compositions are restated here and plugged in with ``oracle.register_host()``:

  * BatchNorm1d / BatchNorm2d forward      nanograd/nn/module.py:296-329, 356-391
  * SGD / Adam / AdamW step                nanograd/optim/optimizer.py:52-55, 79-89, 115-128

They are written in Tensor arithmetic over the oracle namespace (oracle/cpu_namespace.py), so they
reproduce the reference's graph shape (running statistics joining the autograd graph, float64
creep after the first step, ``p.data`` rebinding) and not just its values.
"""
import numpy as np


# ------------------------------------------------------------------ BatchNorm (module.py:296-391)
def _bn_forward(layer, x, view, axes, out_name):
    def normalize(mean, var):
        # module.py:315-329 (1-D divides by sqrt(var+eps)); :377-391 (2-D scales by gamma first)
        if len(view) == 2:
            x_hat = (x - mean.reshape(shape=view)) / (var + layer.eps).sqrt().reshape(shape=view)
            out = layer.weight.reshape(shape=view) * x_hat + layer.bias.reshape(shape=view)
        else:
            x_hat = (x - mean.reshape(shape=view)) * layer.weight.reshape(shape=view)
            out = x_hat / ((var + layer.eps).reshape(shape=view).sqrt()) + layer.bias.reshape(shape=view)
        out.name = out_name
        return out

    if not layer.is_train:
        return normalize(layer.running_mean, layer.running_var)
    count = int(np.prod([x.shape[i] for i in axes]))
    mean = x.mean(axis=axes if len(axes) > 1 else axes[0])
    centered = x - mean.reshape(shape=view)
    var = (centered ** 2).mean(axis=axes if len(axes) > 1 else axes[0])
    # the reference evaluates (x - mean)**2 a second time for the unbiased estimate (module.py:300,366)
    unbiased = ((x - mean.reshape(shape=view)) ** 2).sum(axis=axes if len(axes) > 1 else axes[0]) / (count - 1)
    layer.running_mean = (1. - layer.momentum) * layer.running_mean + layer.momentum * mean
    layer.running_var = (1. - layer.momentum) * layer.running_var + layer.momentum * unbiased
    return normalize(mean, var)


def batchnorm1d_forward(layer, x):
    return _bn_forward(layer, x, [1, -1], (0,), "bn_1d_res")


def batchnorm2d_forward(layer, x):
    return _bn_forward(layer, x, [1, -1, 1, 1], (0, 2, 3), "bn_2d_res")


# ------------------------------------------------------------------ optimizers (optimizer.py)
def sgd_step(opt):
    for i, p in enumerate(opt.params):                                   # optimizer.py:52-55
        opt.momentums[i] = opt.momentums[i] * opt.momentum + opt.lr * p.grad
        p.data = (p - opt.momentums[i]).data


def _moments(opt, i, g):
    opt.exp_avg[i] = opt.beta1 * opt.exp_avg[i] + (1. - opt.beta1) * g
    opt.exp_avg_sq[i] = opt.beta2 * opt.exp_avg_sq[i] + (1. - opt.beta2) * (g ** 2)


def adam_step(opt):
    for i, p in enumerate(opt.params):                                   # optimizer.py:79-89
        _moments(opt, i, p.grad)
        m_hat = opt.exp_avg[i] / (1. - (opt.beta1 ** opt.t))
        v_hat = opt.exp_avg_sq[i] / (1. - (opt.beta2 ** opt.t))
        p.data = (p - opt.lr * m_hat / (v_hat.sqrt() + opt.eps)).data


def adamw_step(opt):
    c1, c2 = 1 - opt.beta1 ** opt.t, 1 - opt.beta2 ** opt.t
    for i, p in enumerate(opt.params):                                   # optimizer.py:115-128
        _moments(opt, i, p.grad)
        denom = (opt.exp_avg_sq[i] / c2).sqrt() + opt.eps
        p.data = (p * (1 - opt.lr * opt.weight_decay) - (opt.lr / c1) * (opt.exp_avg[i] / denom)).data
