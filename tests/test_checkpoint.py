"""Checkpoint / resume: a run that is saved after step 2 and resumed into fresh objects reproduces the
uninterrupted run's steps 3-4 exactly (parameters, BatchNorm running statistics, Adam moments and t)."""
import numpy as np
import pytest

import helpers as H
from nanograd_b200 import checkpoint
from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


def _build(device):
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    np.random.seed(0)
    body = nn.Sequential(nn.Conv2d(3, 8, 3, 1, 1), nn.BatchNorm2d(8), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
                         nn.Linear(8 * 4 * 4, 10))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.body = body

        def forward(self, x):
            return self.body(x).log_softmax()
    net = Net()
    if device == Device.GPU:
        net.gpu()
    # Device.GPU state is float32 throughout, so Adam resumes bit-for-bit; the NumPy oracle drifts to
    # float64 after its first step while a checkpoint stores float32, and Adam would
    # amplify that 1e-8 difference on the zero-gradient conv bias: the oracle device is checked with SGD
    opt = optim.Adam(net.parameters(), lr=1e-2) if device == Device.GPU else optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
    return net, opt, nn.NLLLoss()


def _steps(net, opt, crit, X, Y, device, lo, hi):
    losses = []
    for i in range(lo, hi):
        out = net(Tensor(X[i], device=device))
        opt.zero_grad()
        loss = crit(out, Tensor(Y[i], device=device))
        loss.backward()
        opt.step()
        losses.append(float(np.asarray(loss.numpy() if device == Device.GPU else loss.data).reshape(-1)[0]))
    return losses


def _roundtrip(device, tmp_path):
    rng = np.random.RandomState(1)
    X = rng.randn(4, 6, 3, 8, 8).astype(np.float32)
    Y = rng.randint(0, 10, (4, 6)).astype(np.float32)
    net, opt, crit = _build(device)
    _steps(net, opt, crit, X, Y, device, 0, 2)
    path = str(tmp_path / "ckpt.npz")
    checkpoint.save(path, net, opt)
    want = _steps(net, opt, crit, X, Y, device, 2, 4)
    net2, opt2, crit2 = _build(device)
    checkpoint.load(path, net2, opt2)
    got = _steps(net2, opt2, crit2, X, Y, device, 2, 4)
    tol = 1e-6 if device == Device.GPU else 1e-4
    np.testing.assert_allclose(got, want, rtol=tol, atol=0)
    for a, b in zip(net.parameters(), net2.parameters()):
        np.testing.assert_allclose(b.numpy() if device == Device.GPU else b.data,
                                   a.numpy() if device == Device.GPU else a.data, rtol=tol, atol=tol * 1e-1)


def test_checkpoint_roundtrip_on_the_oracle_device(tmp_path):
    _roundtrip(Device.CPU, tmp_path)


def test_checkpoint_roundtrip_emulated_device(tmp_path):
    H.use_emulated_library()
    _roundtrip(Device.GPU, tmp_path)


@pytest.mark.gpu
def test_checkpoint_roundtrip_on_b200(tmp_path):
    H.use_cuda_library()
    _roundtrip(Device.GPU, tmp_path)


def _swish_net():
    import nanograd_b200.nn.module as nn
    np.random.seed(3)
    return nn.Sequential(nn.Linear(6, 5), nn.Swish(), nn.Linear(5, 2))


def test_checkpoint_with_a_0d_parameter_saved_after_a_step(tmp_path):
    """Swish's learnt beta is 0-d at construction and (1,) after an optimizer step (the reference's
    broadcasting quirk, mirrored by the device step): a save-after-step loads into a fresh model."""
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    H.use_emulated_library()
    net = _swish_net().gpu()
    opt = optim.Adam(net.parameters(), lr=1e-2)
    x = Tensor(np.random.randn(4, 6).astype(np.float32), device=Device.GPU)
    nn.MSELoss()(net(x), Tensor(np.zeros((4, 2), dtype=np.float32), device=Device.GPU)).backward()
    opt.step()
    path = str(tmp_path / "swish.npz")
    checkpoint.save(path, net, opt)
    fresh = _swish_net().gpu()
    fresh_opt = optim.Adam(fresh.parameters(), lr=1e-2)
    checkpoint.load(path, fresh, fresh_opt)
    for a, b in zip(net.parameters(), fresh.parameters()):
        np.testing.assert_array_equal(a.numpy().reshape(-1), b.numpy().reshape(-1))
    assert fresh_opt.t == 1


def test_checkpoint_refuses_a_reordered_model(tmp_path):
    import nanograd_b200.nn.module as nn
    H.use_emulated_library()

    class A(nn.Module):
        def __init__(self):
            super().__init__()
            self.first = nn.Linear(4, 4)
            self.second = nn.Linear(4, 4)

    class B(nn.Module):
        def __init__(self):
            super().__init__()
            self.second = nn.Linear(4, 4)
            self.first = nn.Linear(4, 4)

    path = str(tmp_path / "order.npz")
    checkpoint.save(path, A())
    with pytest.raises(Exception, match="parameter order differs"):
        checkpoint.load(path, B())
    checkpoint.load(path, A())
