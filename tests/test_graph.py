"""Captured steps (nanograd_b200.graph.GraphedStep): replaying the CUDA graph of a README-CNN training step
gives bit-for-bit the parameters and losses of the eager loop, on changing batches; memory allocated while
capturing stays with the graph; optimizers with per-step host scalars are refused."""
import numpy as np
import pytest

import helpers as H
from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


def _setup(opt_kind="sgd"):
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    np.random.seed(0)
    model = H.build_mnist_cnn(nn).gpu()
    params = list(model.parameters())
    opt = optim.SGD(params, lr=0.01, momentum=0.9) if opt_kind == "sgd" else optim.Adam(params)
    crit = nn.NLLLoss()

    def step(xb, yb):
        out = model(xb)
        opt.zero_grad()
        loss = crit(out, yb)
        loss.backward()
        opt.step()
        return loss
    return model, step


def _batches(n, batch=32):
    rng = np.random.RandomState(3)
    return rng.rand(n, batch, 784).astype(np.float32), rng.randint(0, 10, (n, batch)).astype(np.float32)


@pytest.mark.gpu
def test_graphed_step_matches_eager_bit_for_bit():
    from nanograd_b200 import backend
    from nanograd_b200.graph import GraphedStep
    H.use_cuda_library()
    X, Y = _batches(6)
    # eager: 3 steps on batch 0 (what GraphedStep's 2 warm-up steps + first replay do), then batches 1..5
    model_e, step_e = _setup()
    losses_e = []
    for i in [0, 0, 0, 1, 2, 3, 4, 5]:
        losses_e.append(float(step_e(Tensor(X[i], device=Device.GPU), Tensor(Y[i], device=Device.GPU)).numpy()[0]))
    model_g, step_g = _setup()
    live_before = backend.mem_stats()["live_bytes"]
    fast = GraphedStep(step_g, X[0], Y[0], warmup=2)
    assert fast.kernel_nodes >= 20 and fast.private_bytes > 0
    losses_g = [float(fast.outputs.numpy()[0])]
    launches = backend.launch_count()
    for i in range(1, 6):
        losses_g.append(float(fast(X[i], Y[i]).numpy()[0]))
    assert backend.launch_count() - launches == 5 * fast.kernel_nodes
    assert losses_g == losses_e[2:], (losses_g, losses_e[2:])
    for a, b in zip(model_g.parameters(), model_e.parameters()):
        H.assert_exact(a.numpy(), b.numpy(), "parameter after graphed vs eager steps")
    # eager work after the capture does not disturb the graph's memory (and vice versa)
    scratch = [Tensor(np.random.rand(64, 784).astype(np.float32), device=Device.GPU) for _ in range(8)]
    again = float(fast(X[5], Y[5]).numpy()[0])
    assert np.isfinite(again)
    del scratch
    fast.close()
    assert backend.mem_stats()["live_bytes"] >= 0 and live_before >= 0


@pytest.mark.gpu
def test_graphed_step_refuses_adam_and_wrong_shapes():
    from nanograd_b200.graph import GraphedStep
    H.use_cuda_library()
    X, Y = _batches(2)
    _, step_adam = _setup("adam")
    with pytest.raises(Exception, match="bias correction"):
        GraphedStep(step_adam, X[0], Y[0])
    _, step_sgd = _setup()
    fast = GraphedStep(step_sgd, X[0], Y[0])
    with pytest.raises(Exception, match="captured for input shape"):
        fast(X[0][:16], Y[0][:16])
    # the library is usable eagerly after a refused / closed capture
    fast.close()
    _, step2 = _setup()
    assert np.isfinite(float(step2(Tensor(X[1], device=Device.GPU), Tensor(Y[1], device=Device.GPU)).numpy()[0]))


def test_graph_capture_is_refused_by_the_emulation_build():
    from nanograd_b200 import backend
    H.use_emulated_library()
    with pytest.raises(Exception, match="not available in the emulation build"):
        backend.lib().ng_graph_begin()
    assert backend.graph_capturing() is False
