"""the layers the reference composes from primitives and this backend fuses —
nn.LeakyReLU (module.py:744-772), nn.Swish (:818-841), nn.Mish (:844-864), nn.MSELoss (:689-712) —
on Device.GPU against the same module on Device.CPU (the NumPy oracle running the reference's
composition), forward and every gradient at the FP32 bar (rtol 1e-4, scale-relative).
The inputs contain exact zeros (LeakyReLU's gradient there is 1 + alpha in the reference) and
values out to +-15 (softplus / sigmoid tails)."""
import numpy as np
import pytest

import helpers as H


def _inputs(shape, seed):
    rng = np.random.RandomState(seed)
    x = rng.normal(0, 3, shape).astype(np.float32)
    flat = x.reshape(-1)
    flat[::7] = 0.0
    flat[3::31] = 15.0
    flat[5::37] = -15.0
    return x


def _run_layers():
    import nanograd_b200.nn.module as nn
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor
    for name in ("leakyrelu", "swish", "mish", "mseloss"):
        assert name in Tensor.ops[Device.GPU], f"fused op {name} is not registered on the device namespace"
        assert name not in Tensor.ops[Device.CPU]
    x = _inputs((33, 45), 0)
    up = np.random.RandomState(1).normal(0, 1, x.shape).astype(np.float32)

    def both(make_layer, extra_params=()):
        res = []
        for device in (Device.CPU, Device.GPU):
            layer = make_layer()
            if device == Device.GPU:
                layer.gpu()
            t = Tensor(x.copy(), requires_grad=True, device=device)
            out = layer(t)
            (out * Tensor(up.copy(), device=device)).sum().backward()
            res.append([out.numpy(), t.grad.numpy()] + [getattr(layer, p).grad.numpy() for p in extra_params])
        for i, (c, g) in enumerate(zip(*res)):
            H.assert_close(g, c, H.RTOL_FP32, f"{type(make_layer()).__name__} result {i}")
        return res

    lk = both(lambda: nn.LeakyReLU(0.2))
    zero = (x == 0)
    assert zero.any() and np.allclose(lk[1][1][zero], 1.2 * up[zero], rtol=1e-6), "LeakyReLU'(0) must be 1 + alpha"
    both(lambda: nn.Mish())
    both(lambda: nn.Swish(1.3), extra_params=("beta",))

    # MSELoss: both arguments ask for a gradient
    t_np = np.random.RandomState(2).normal(0, 2, x.shape).astype(np.float32)
    res = []
    for device in (Device.CPU, Device.GPU):
        p = Tensor(x.copy(), requires_grad=True, device=device)
        t = Tensor(t_np.copy(), requires_grad=True, device=device)
        loss = nn.MSELoss()(p, t)
        (loss * 3.0).backward()
        res.append([loss.numpy().reshape(-1), p.grad.numpy(), t.grad.numpy()])
    for i, (c, g) in enumerate(zip(*res)):
        H.assert_close(g, c, H.RTOL_FP32, f"MSELoss result {i}")


def test_fused_layers_match_the_oracle_emulated():
    H.use_emulated_library()
    _run_layers()


@pytest.mark.gpu
def test_fused_layers_match_the_oracle_on_b200():
    H.use_cuda_library()
    _run_layers()


@pytest.mark.gpu
def test_fused_layers_at_activation_scale_on_b200():
    """Size-independent properties at a BASELINE-sized activation (512 x 64 x 32 x 32 = 33.5 M values):
    LeakyReLU(alpha=1) is the identity, Swish(beta=0) is x / 2, Mish is odd-asymmetric but bounded by
    |x|, MSELoss(x, x) is exactly 0 and MSELoss(x, 0) is mean(x^2)."""
    import nanograd_b200.nn.module as nn
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor
    H.use_cuda_library()
    x_np = np.random.RandomState(0).normal(0, 2, (512, 64, 32, 32)).astype(np.float32)
    x = Tensor(x_np, device=Device.GPU)
    H.assert_exact(nn.LeakyReLU(1.0).gpu()(x).numpy(), x_np, "LeakyReLU(alpha=1)")
    H.assert_exact(nn.Swish(0.0).gpu()(x).numpy(), x_np / 2, "Swish(beta=0)")
    m = nn.Mish().gpu()(x).numpy()
    assert np.all(np.abs(m) <= np.abs(x_np) + 1e-6) and np.all(m >= -0.31)
    assert float(nn.MSELoss()(x, x).numpy().reshape(-1)[0]) == 0.0
    zero = Tensor(np.zeros_like(x_np), device=Device.GPU)
    H.assert_close(nn.MSELoss()(x, zero).numpy().reshape(-1), [np.mean(x_np.astype(np.float64) ** 2)], H.RTOL_FP32, "mean(x^2)")
