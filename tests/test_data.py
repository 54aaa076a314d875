"""Input pipeline, device-side accuracy and the train / evaluate loops."""
import gzip
import os

import numpy as np
import pytest

import helpers as H
from nanograd_b200 import data
from nanograd_b200.device import Device


def _write_idx(tmp_path, n_train=7, n_test=3):
    rng = np.random.RandomState(0)
    imgs = {k: rng.randint(0, 256, (n, 784)).astype(np.uint8) for k, n in (("train", n_train), ("test", n_test))}
    labs = {k: rng.randint(0, 10, n).astype(np.uint8) for k, n in (("train", n_train), ("test", n_test))}
    paths = []
    for k in ("train", "test"):
        for kind, header, payload in (("images", 16, imgs[k]), ("labels", 8, labs[k])):
            path = os.path.join(tmp_path, f"{k}-{kind}.gz")
            with open(path, "wb") as f:
                f.write(gzip.compress(bytes(header) + payload.tobytes()))
            paths.append(path)
    return paths, imgs, labs


def test_load_mnist_idx_matches_the_reference_reader(tmp_path):
    paths, imgs, labs = _write_idx(str(tmp_path))
    xtr, ytr, xte, yte = data.load_mnist_idx(paths)
    assert xtr.shape == (7, 784) and xte.shape == (3, 784) and xtr.dtype == np.float32
    np.testing.assert_array_equal(ytr, labs["train"])
    np.testing.assert_array_equal(yte, labs["test"])
    np.testing.assert_allclose(xtr, imgs["train"].astype(np.float32) / 255.0, rtol=0, atol=1e-7)


def _toy(n=96, d=20, classes=4, seed=3):
    rng = np.random.RandomState(seed)
    centers = rng.randn(classes, d) * 3
    y = rng.randint(0, classes, n)
    x = (centers[y] + rng.randn(n, d)).astype(np.float32)
    return x, y.astype(np.float32)


def test_cpu_loader_draws_the_reference_batches():
    x, y = _toy()
    np.random.seed(5)
    want = [np.random.randint(0, x.shape[0], size=(8,)) for _ in range(4)]
    np.random.seed(5)
    got = list(data.BatchLoader(x, y, 8, 4, device=Device.CPU))
    for idx, (xb, yb) in zip(want, got):
        np.testing.assert_array_equal(np.asarray(xb.data), x[idx])
        np.testing.assert_array_equal(np.asarray(yb.data), y[idx])


def _device_loader_checks():
    x, y = _toy()
    np.random.seed(11)
    want = [np.random.randint(0, x.shape[0], size=(16,)) for _ in range(5)]
    np.random.seed(11)
    for idx, (xb, yb) in zip(want, data.BatchLoader(x, y, 16, 5, device=Device.GPU)):
        assert xb.device == Device.GPU
        np.testing.assert_array_equal(xb.numpy(), x[idx])   # prefetched through the copy stream
        np.testing.assert_array_equal(yb.numpy(), y[idx])
    # sequential evaluation batches with a ragged tail
    seen = []
    n = x.shape[0]
    ld = data.BatchLoader(x, y, 40, 3, device=Device.GPU, indices=lambda i: np.arange(40 * i, min(40 * (i + 1), n)))
    for xb, yb in ld:
        seen.append(xb.numpy())
    np.testing.assert_array_equal(np.concatenate(seen), x)
    # device-side accuracy == NumPy argmax accuracy (first maximum wins on ties)
    from nanograd_b200.tensor import Tensor
    rng = np.random.RandomState(2)
    logits = rng.randn(50, 10).astype(np.float32)
    logits[3, :] = 1.0
    labels = rng.randint(0, 10, 50).astype(np.float32)
    acc = data.batch_accuracy(Tensor(logits, device=Device.GPU), Tensor(labels, device=Device.GPU))
    assert acc == pytest.approx(float((np.argmax(logits, -1) == labels).mean()), abs=1e-7)


def _train_eval_checks(device):
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    x, y = _toy(n=256)
    np.random.seed(0)
    model = nn.Sequential(nn.Linear(20, 32), nn.ReLU(), nn.Linear(32, 4))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.body = model

        def forward(self, t):
            return self.body(t).log_softmax()

    net = Net()
    if device == Device.GPU:
        net.gpu()
    opt = optim.SGD(net.parameters(), lr=0.5, momentum=0.9)
    losses, accs = data.train(net, x, y, opt, steps=60, batch_size=32, device=device, log_every=10)
    assert losses[-1] < 0.5 * losses[0]
    assert data.evaluate(net, x, y, batch_size=100, device=device) > 0.9


def test_train_and_evaluate_on_the_oracle_device():
    _train_eval_checks(Device.CPU)


def test_device_loader_and_accuracy_emulated():
    H.use_emulated_library()
    _device_loader_checks()


@pytest.mark.gpu
def test_device_loader_accuracy_and_training_on_b200():
    H.use_cuda_library()
    _device_loader_checks()
    _train_eval_checks(Device.GPU)


def test_graph_visualiser_draws_forward_and_backward_graphs():
    """north-star: 'the graph visualiser keeps working' (at HEAD the reference's is broken: children are
    never populated and plot_backward raises).  Runs on the oracle device."""
    from nanograd_b200.tensor import Tensor
    a = Tensor(np.ones((2, 3), dtype=np.float32), requires_grad=True, name="a")
    b = Tensor(np.full((2, 3), 2.0, dtype=np.float32), requires_grad=True, name="b")
    f = ((a * b) + a).sum()
    f.backward()
    fwd = f.plot_forward()
    bwd = f.plot_backward()
    for g in (fwd, bwd):
        src = g.source if isinstance(g.source, str) else g.source()
        assert src.count("->") >= 3, src      # edges for mul, add, sum


def test_graph_visualiser_on_device_tensors(device_lib):
    """The same plots for a model on Device.GPU (host emulation in the CPU suite, the B200 under -m gpu):
    node labels come from ``.shape`` / ``.op`` only, nothing is copied off the device, and a no-grad
    inference chain keeps no edges (and therefore pins no device buffers)."""
    import nanograd_b200.nn.module as nn
    from nanograd_b200 import backend
    from nanograd_b200.tensor import Tensor
    np.random.seed(0)
    model = H.build_mnist_cnn(nn).gpu()
    x = Tensor(np.random.rand(4, 784).astype(np.float32), device=Device.GPU)
    y = Tensor(np.random.randint(0, 10, 4).astype(np.float32), device=Device.GPU)
    loss = nn.NLLLoss()(model(x), y)
    loss.backward()
    launches = backend.launch_count()
    fwd, bwd = loss.plot_forward(), loss.plot_backward()
    assert backend.launch_count() == launches            # drawing launches nothing on the device
    for g in (fwd, bwd):
        src = g.source if isinstance(g.source, str) else g.source()
        assert src.count("->") >= 10, src                # conv, relu, pool, conv, relu, pool, reshape, linear, ...
    ops = {n.op for n in _walk(loss)}
    assert {"conv2dbias", "maxpool2d", "linear", "logsoftmax", "nllloss"} <= ops, ops
    # inference chain: no requires_grad inputs -> no children retained
    frozen = Tensor(np.random.rand(2, 5).astype(np.float32), device=Device.GPU)
    out = (frozen * 2.0).relu().sum()
    assert out.children == [] and out.ctx is None


def _walk(root):
    seen, stack = [], [root]
    while stack:
        n = stack.pop()
        if any(n is s for s in seen):
            continue
        seen.append(n)
        stack.extend(n.children)
    return seen
