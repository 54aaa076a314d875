"""Host-side optimizer behaviour on Device.GPU: state created before ``model.gpu()``
is moved on the first step (the reference's flow, tests/test_mnist.py:95-96 + helpers.py:287-300), and
``Tensor.copy()`` snapshots values although the device step updates parameters in place."""
import numpy as np
import pytest

import helpers as H
from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


@pytest.mark.parametrize("kind", ["sgd", "adam", "adamw"])
def test_optimizer_built_before_gpu_move(kind, device_lib):
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    np.random.seed(0)
    model = nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))
    params = list(model.parameters())
    opt = {"sgd": lambda: optim.SGD(params, lr=0.01, momentum=0.9), "adam": lambda: optim.Adam(params),
           "adamw": lambda: optim.AdamW(params)}[kind]()
    model.gpu()   # after the optimizer exists: its state tensors are still host arrays
    x = Tensor(np.random.randn(4, 6).astype(np.float32), device=Device.GPU)
    for _ in range(2):
        opt.zero_grad()
        model(x).sum().backward()
        opt.step()
    state = opt.momentums if kind == "sgd" else opt.exp_avg
    assert all(s.device == Device.GPU for s in state)
    assert all(np.isfinite(p.numpy()).all() for p in params)


def test_train_helper_with_optimizer_built_on_the_host(device_lib):
    from nanograd_b200 import data
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    np.random.seed(1)
    model = H.build_mnist_cnn(nn)
    opt = optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    X = np.random.rand(16, 784).astype(np.float32)
    Y = np.random.randint(0, 10, 16).astype(np.float32)
    data.train(model, X, Y, opt, steps=2, batch_size=4, device=Device.GPU)
    assert all(m.device == Device.GPU for m in opt.momentums)


def test_copy_is_a_snapshot_under_the_in_place_step(device_lib):
    import nanograd_b200.optim.optimizer as optim
    w = Tensor(np.ones((3, 4), dtype=np.float32), requires_grad=True, is_parameter=True, device=Device.GPU)
    before = w.copy()
    w.grad = Tensor(np.full((3, 4), 2.0, dtype=np.float32), device=Device.GPU)
    optim.SGD([w], lr=0.5, momentum=0).step()
    np.testing.assert_array_equal(before.numpy(), np.ones((3, 4), dtype=np.float32))
    np.testing.assert_array_equal(w.numpy(), np.zeros((3, 4), dtype=np.float32))
