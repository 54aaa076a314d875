"""Edge shapes of the generic operators on the emulated build of the kernel sources (tests/emu), which enforces
the launch limits of a real GPU (no zero-sized grid, grid.y / grid.z <= 65535, <= 227 KB of dynamic shared
memory — tests/emu/ng_emu.cpp aborts otherwise): leading dimensions beyond 65535 (a batch or row count that a
kernel must not map to grid.y), empty tensors, single rows.  Values are compared with the NumPy oracle on the
same seeded inputs (ops_cpu.py semantics), gradients included.
"""
import numpy as np
import pytest

import helpers as H

BIG = 70001   # > 65535 and odd


@pytest.fixture(autouse=True)
def _emu():
    H.use_emulated_library()


def test_large_leading_dimension_elementwise_reduce_softmax():
    H.check_op([(BIG, 3)], lambda x: (x.relu() * 2.0 + 1.0).sum(axis=1), loc=0.0)
    H.check_op([(BIG, 3)], lambda x: x.sum(axis=0), loc=0.0)
    H.check_op([(BIG, 3)], lambda x: x.max(axis=1), loc=0.0)


def test_large_row_count_log_softmax():
    H.check_op([(BIG, 10)], lambda x: x.log_softmax(), loc=0.0)


def test_large_leading_dimension_transpose_slice_matmul():
    H.check_op([(BIG, 3)], lambda x: x.transpose(), loc=0.0)
    H.check_op([(BIG, 3)], lambda x: x[10:BIG - 1000, 1:3], loc=0.0)
    H.check_op([(BIG, 3), (3, 2)], lambda x, w: x @ w, loc=0.0)


def test_large_batch_pooling():
    H.check_op([(BIG, 1, 2, 2)], lambda x: x.max_pool2d(), loc=0.0)
    H.check_op([(BIG, 1, 2, 2)], lambda x: x.avg_pool2d(), loc=0.0)


def test_large_batch_conv():
    H.check_op([(BIG, 1, 1, 1), (2, 1, 1, 1)], lambda x, w: x.conv2d(w, stride=1), loc=0.0)


def test_large_batch_batchnorm_and_loss():
    from nanograd_b200.device import Device
    fw_g, fw_c = H.device_framework(Device.GPU), H.device_framework(Device.CPU)
    rng = np.random.RandomState(0)
    X = rng.randn(BIG, 2, 1, 1).astype(np.float32)
    Y = rng.randint(0, 2, BIG).astype(np.float32)
    outs = []
    for fw in (fw_c, fw_g):
        np.random.seed(0)
        bn = fw.place(fw.nn.BatchNorm2d(2))
        x = fw.tensor(X, requires_grad=True)
        logits = bn(x).reshape(shape=[BIG, 2])
        loss = fw.nn.NLLLoss()(logits.log_softmax(), fw.tensor(Y))
        loss.backward()
        outs.append((fw.numpy(loss), fw.numpy(x.grad), fw.numpy(bn.weight.grad)))
    for got, want, what in zip(outs[1], outs[0], ("loss", "dx", "dgamma")):
        H.assert_close(got, want, H.RTOL_FP32, what)


def test_empty_and_single_row_tensors():
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor
    e = Tensor(np.zeros((0, 4), dtype=np.float32), device=Device.GPU, requires_grad=True)
    assert (e.relu() * 2.0 + 1.0).numpy().shape == (0, 4)          # nothing to launch
    np.testing.assert_array_equal(e.sum(axis=0).numpy(), np.zeros(4, np.float32))   # np.sum over an empty axis
    with pytest.raises(ValueError):
        e.max(axis=0)                                               # np.amax has no identity
    H.check_op([(1, 5)], lambda x: x.log_softmax(), loc=0.0)
    H.check_op([(1, 1, 4, 4)], lambda x: x.max_pool2d(), loc=0.0)
