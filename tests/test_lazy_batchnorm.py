"""BatchNorm2d as "statistics now, element-wise pass on demand" (nn/buffer.py LazyBuffer).

In a conv -> BatchNorm2d -> ReLU -> conv chain the normalised activation (and, in the backward, BatchNorm's input
gradient) is read by nobody but the staging pass of the next convolution's FP16x3 tensor-core operands, so the
op returns a recipe and the convolution stages straight from it (ng_conv2d_stage_f16_bn / _bnbwd).  Checked here:
  * a recipe that IS materialised (pooling, numpy(), any generic op reads .ptr) is bit-identical to the one-call
    kernels (module.py:356-391's values) — also under the host emulation of the CPU suite;
  * the magnitudes the statistics passes report bound the tensors they describe (exactly, for the forward);
  * the fused staging copies equal the staging copy of the materialised tensor to the 2^-21 of the FP16 pairs,
    the channel sums (conv bias gradient, module.py:476) included.
Whole-model parity of the fused path: the vgg_w64_* cases of tests/test_device_parity.py (fp32 and f16x3 modes).
"""
import numpy as np
import pytest

import helpers as H


def _setup(n, c, h, w, seed):
    from nanograd_b200.nn.buffer import DeviceBuffer
    rng = np.random.RandomState(seed)
    x = (rng.randn(n, c, h, w) * rng.uniform(0.2, 3.0, (1, c, 1, 1)) + rng.uniform(-2, 2, (1, c, 1, 1))).astype(np.float32)
    gamma = rng.uniform(-1.5, 1.5, c).astype(np.float32)
    beta = rng.uniform(-0.5, 0.5, c).astype(np.float32)
    dy = (rng.randn(n, c, h, w) * np.exp(rng.randn(n, c, h, w))).astype(np.float32)
    dev = {k: DeviceBuffer(v.shape, hostbuf=v) for k, v in dict(x=x, gamma=gamma, beta=beta, dy=dy).items()}
    return dict(x=x, gamma=gamma, beta=beta, dy=dy), dev


def _one_call_and_split(lib, dev, n, c, hw, act):
    """(y, mean, invstd, dx, dgamma, dbeta) from the one-call kernels and from the split kernels."""
    from nanograd_b200.nn.buffer import DeviceBuffer
    B = lambda *s: DeviceBuffer(s)
    y1, m1, i1, dx1, dg1, db1 = B(n, c, hw), B(c), B(c), B(n, c, hw), B(c), B(c)
    lib.ng_bn2d_act_fwd_train(y1.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, 0, 0, m1.ptr, i1.ptr, n, c, hw,
                              1e-5, 0.1, act)
    lib.ng_bn2d_act_bwd(dx1.ptr, dg1.ptr, db1.ptr, dev["dy"].ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr,
                        m1.ptr, i1.ptr, n, c, hw, act)
    y2, m2, i2, dx2, dg2, db2 = B(n, c, hw), B(c), B(c), B(n, c, hw), B(c), B(c)
    absmax, sums, bound = B(1), B(2 * c + 1), B(1)
    lib.ng_bn2d_act_fwd_stats(dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, 0, 0, m2.ptr, i2.ptr, absmax.ptr, 0, n, c, hw,
                              1e-5, 0.1, act)
    lib.ng_bn2d_act_apply(y2.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, m2.ptr, i2.ptr, n, c, hw, act)
    lib.ng_bn2d_act_bwd_sums(dg2.ptr, db2.ptr, sums.ptr, bound.ptr, dev["dy"].ptr, dev["x"].ptr, dev["gamma"].ptr,
                             dev["beta"].ptr, m2.ptr, i2.ptr, n, c, hw, act)
    lib.ng_bn2d_act_bwd_dx(dx2.ptr, dev["dy"].ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, m2.ptr, i2.ptr,
                           sums.ptr, n, c, hw, act)
    one = [b.numpy() for b in (y1, m1, i1, dx1, dg1, db1)]
    split = [b.numpy() for b in (y2, m2, i2, dx2, dg2, db2)]
    return one, split, (absmax, sums, bound, m2, i2)


def _check_split(act, shape):
    from nanograd_b200 import backend as B
    n, c, h, w = shape
    host, dev = _setup(n, c, h, w, seed=c + act)
    one, split, (absmax, _, bound, _, _) = _one_call_and_split(B.lib(), dev, n, c, h * w, act)
    for name, a, b in zip(("y", "mean", "invstd", "dx", "dgamma", "dbeta"), one, split):
        H.assert_exact(b, a, f"{name} (split kernels vs one-call kernels, act={act})")
    y, dx = one[0], one[3]
    amax = float(absmax.numpy().view(np.float32)[0])
    assert float(np.abs(y).max()) <= amax <= float(np.abs(y).max()) * 1.001 + 1e-30, (amax, float(np.abs(y).max()))
    bnd = float(bound.numpy().view(np.float32)[0])
    assert float(np.abs(dx).max()) <= bnd <= 16 * float(np.abs(dx).max()), (bnd, float(np.abs(dx).max()))


@pytest.mark.parametrize("act", [0, 1])
def test_split_batchnorm_kernels_match_one_call_kernels_emulated(act):
    H.use_emulated_library()
    _check_split(act, (3, 8, 5, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("shape", [(3, 8, 5, 4), (16, 64, 16, 16), (7, 192, 9, 5), (64, 128, 32, 32)],
                         ids=lambda s: "x".join(map(str, s)))
def test_split_batchnorm_kernels_match_one_call_kernels(shape, act):
    H.use_cuda_library()
    _check_split(act, shape)


def _decode(staged, n, c, hw):
    raw = staged.numpy()
    total = n * c * hw
    inv_scale = float(raw[total])
    halves = raw[:total].view(np.float16).reshape(n, hw, c // 64, 2, 64).astype(np.float64)
    vals = (halves[:, :, :, 0, :] + halves[:, :, :, 1, :]).reshape(n, hw, c) * inv_scale
    return vals.transpose(0, 2, 1), inv_scale   # [n][c][hw]


@pytest.mark.gpu
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("shape", [(16, 64, 16, 16), (7, 192, 9, 5), (32, 128, 32, 32)], ids=lambda s: "x".join(map(str, s)))
def test_fused_staging_equals_staging_of_the_materialised_tensor(shape, act):
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    n, c, h, w = shape
    hw = h * w
    host, dev = _setup(n, c, h, w, seed=7 * c + act)
    one, _, (absmax, sums, bound, mean, invstd) = _one_call_and_split(lib, dev, n, c, hw, act)
    y, dx = one[0].astype(np.float64), one[3].astype(np.float64)
    staged = DeviceBuffer((n * c * hw + 16,))
    lib.ng_conv2d_stage_f16_bn(staged.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, mean.ptr, invstd.ptr,
                               absmax.ptr, n, c, hw, act)
    got, inv_scale = _decode(staged, n, c, hw)
    ymax = np.abs(y).max()
    assert 2.0 ** 13 <= ymax / inv_scale < 2.0 ** 15, "scale puts the largest magnitude just under 2^15"
    err = np.abs(got - y)
    assert np.all(err <= np.maximum(np.abs(y) * 2.0 ** -21, ymax * 2.0 ** -38)), float(err.max() / ymax)

    staged2, chan = DeviceBuffer((n * c * hw + 16,)), DeviceBuffer((c,))
    lib.ng_conv2d_stage_f16_bnbwd(staged2.ptr, chan.ptr, dev["dy"].ptr, dev["x"].ptr, dev["gamma"].ptr,
                                  dev["beta"].ptr if act else 0, mean.ptr, invstd.ptr, sums.ptr, bound.ptr, n, c, hw, act)
    got, inv_scale = _decode(staged2, n, c, hw)
    dmax = np.abs(dx).max()
    assert 2.0 ** 10 <= dmax / inv_scale < 2.0 ** 15, "the bound costs at most a few binades of FP16 range"
    # (the fused kernel re-evaluates k0 (dy - k1 - (x - mean) k2) in float32: where the terms cancel, the two
    #  evaluations may contract their multiply-adds differently, hence the absolute floor of 2^-20 max|dx|)
    err = np.abs(got - dx)
    assert np.all(err <= np.maximum(np.abs(dx) * 2.0 ** -21, dmax * 2.0 ** -20)), float(err.max() / dmax)
    # (the channel sums of BatchNorm's dx vanish analytically: compare on the scale of sum |dx|)
    np.testing.assert_allclose(chan.numpy(), dx.sum(axis=(0, 2)), rtol=0, atol=1e-5 * float(np.abs(dx).sum(axis=(0, 2)).max()),
                               err_msg="channel sums of the fused dx staging")


@pytest.mark.gpu
def test_lazy_outputs_are_only_written_when_somebody_reads_them():
    """conv -> BatchNorm2d -> ReLU -> conv on the FP16x3 engine: the normalised activation and BatchNorm's input
    gradient stay recipes (never allocated); a max-pool consumer materialises its input."""
    from nanograd_b200 import backend as B
    from nanograd_b200.device import Device
    from nanograd_b200.nn.buffer import LazyBuffer
    import nanograd_b200.nn.module as nn
    from nanograd_b200.tensor import Tensor
    H.use_cuda_library()
    B.set_math_mode("f16x3")
    try:
        np.random.seed(3)
        c1, bn, c2, pool = nn.Conv2d(64, 64, 3, 1, 1), nn.BatchNorm2d(64), nn.Conv2d(64, 128, 3, 1, 1), nn.MaxPool2d(2)
        for m in (c1, bn, c2):
            m.gpu()
        x = Tensor(np.random.randn(4, 64, 8, 8).astype(np.float32), device=Device.GPU)
        a = c1(x)
        z = bn.forward(a, relu=True)
        assert isinstance(z.data, LazyBuffer) and z.data.pending
        y = c2(z)
        assert z.data.pending, "the second convolution staged its operand from the recipe"
        p = pool(y)
        loss = (p * p).sum()
        loss.backward()
        assert isinstance(a.grad.data, LazyBuffer) and a.grad.data.pending, "dx of BatchNorm stayed a recipe"
        assert z.data.pending
        # reading the values materialises them, and they are BatchNorm + ReLU of a
        zv, av = z.numpy(), a.numpy().astype(np.float64)
        mu, var = av.mean(axis=(0, 2, 3), keepdims=True), av.var(axis=(0, 2, 3), keepdims=True)
        want = np.maximum((av - mu) / np.sqrt(var + 1e-5) * bn.weight.numpy().reshape(1, -1, 1, 1)
                          + bn.bias.numpy().reshape(1, -1, 1, 1), 0)
        H.assert_close(zv, want, H.RTOL_FP32, "materialised BatchNorm + ReLU output")
        assert not z.data.pending
    finally:
        B.set_math_mode("fp32")


def _check_fused_maxpool(shape, act):
    """Max pooling straight from the BatchNorm recipe: outputs and tie-splitting gradients bit-identical to pooling
    the materialised tensor (the window maximum and the equality mask see the same float32 values)."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    lib = B.lib()
    n, c, h, w = shape
    hw = h * w
    host, dev = _setup(n, c, h, w, seed=11 * c + act)
    # tie-rich input: quantised values make equal window maxima common (and ReLU zeros tie by themselves)
    xq = np.round(host["x"] * 2) / 2
    dev["x"] = DeviceBuffer(xq.shape, hostbuf=xq.astype(np.float32))
    y, mean, invstd, absmax = DeviceBuffer((n, c, h, w)), DeviceBuffer((c,)), DeviceBuffer((c,)), DeviceBuffer((1,))
    lib.ng_bn2d_act_fwd_stats(dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, 0, 0, mean.ptr, invstd.ptr, absmax.ptr, 0,
                              n, c, hw, 1e-5, 0.1, act)
    lib.ng_bn2d_act_apply(y.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, mean.ptr, invstd.ptr, n, c, hw, act)
    oh, ow = h // 2, w // 2
    g = np.random.RandomState(5).randn(n, c, oh, ow).astype(np.float32)
    gd = DeviceBuffer(g.shape, hostbuf=g)
    p1, d1 = DeviceBuffer((n, c, oh, ow)), DeviceBuffer((n, c, h, w))
    lib.ng_maxpool2d_fwd(p1.ptr, y.ptr, n * c, h, w, 2, 2)
    lib.ng_maxpool2d_bwd(d1.ptr, gd.ptr, y.ptr, p1.ptr, n * c, h, w, 2, 2)
    p2, d2 = DeviceBuffer((n, c, oh, ow)), DeviceBuffer((n, c, h, w))
    lib.ng_maxpool2d_fwd_bn(p2.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, mean.ptr, invstd.ptr, n * c, c, h, w, 2, 2, act)
    lib.ng_maxpool2d_bwd_bn(d2.ptr, gd.ptr, dev["x"].ptr, p2.ptr, dev["gamma"].ptr, dev["beta"].ptr, mean.ptr, invstd.ptr,
                            n * c, c, h, w, 2, 2, act)
    H.assert_exact(p2.numpy(), p1.numpy(), "fused max-pool output")
    H.assert_exact(d2.numpy(), d1.numpy(), "fused max-pool input gradient")
    assert (d1.numpy() != 0).sum() > n * c * oh * ow, "the data has ties (gradient split among several window elements)"


@pytest.mark.parametrize("act", [0, 1])
def test_fused_maxpool_matches_pooling_the_materialised_tensor_emulated(act):
    H.use_emulated_library()
    _check_fused_maxpool((2, 6, 6, 8), act)


@pytest.mark.gpu
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("shape", [(2, 6, 6, 8), (16, 64, 32, 32), (5, 128, 7, 9)], ids=lambda s: "x".join(map(str, s)))
def test_fused_maxpool_matches_pooling_the_materialised_tensor(shape, act):
    H.use_cuda_library()
    _check_fused_maxpool(shape, act)


def _pooled_setup(shape, act, seed):
    """A conv -> BatchNorm(+ReLU) -> MaxPool2d(2) backward: x (layer input), the pooled maxima, their gradient,
    and the materialised chain (pool backward -> BatchNorm sums -> dx) to compare the fused kernels with."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    lib = B.lib()
    n, c, h, w = shape
    hw = h * w
    host, dev = _setup(n, c, h, w, seed=seed)
    xq = np.round(host["x"] * 2) / 2   # tie-rich
    dev["x"] = DeviceBuffer(xq.shape, hostbuf=xq.astype(np.float32))
    mean, invstd, absmax = DeviceBuffer((c,)), DeviceBuffer((c,)), DeviceBuffer((1,))
    lib.ng_bn2d_act_fwd_stats(dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, 0, 0, mean.ptr, invstd.ptr, absmax.ptr, 0,
                              n, c, hw, 1e-5, 0.1, act)
    oh, ow = h // 2, w // 2
    pooled = DeviceBuffer((n, c, oh, ow))
    lib.ng_maxpool2d_fwd_bn(pooled.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, mean.ptr, invstd.ptr, n * c, c, h, w, 2, 2, act)
    g = (np.random.RandomState(seed + 1).randn(n, c, oh, ow) * 0.1).astype(np.float32)
    dpool = DeviceBuffer(g.shape, hostbuf=g)
    # materialised chain
    dy = DeviceBuffer((n, c, h, w))
    lib.ng_maxpool2d_bwd_bn(dy.ptr, dpool.ptr, dev["x"].ptr, pooled.ptr, dev["gamma"].ptr, dev["beta"].ptr, mean.ptr, invstd.ptr,
                            n * c, c, h, w, 2, 2, act)
    dg1, db1, sums1, bound1, dx1 = DeviceBuffer((c,)), DeviceBuffer((c,)), DeviceBuffer((2 * c + 1,)), DeviceBuffer((1,)), DeviceBuffer((n, c, h, w))
    lib.ng_bn2d_act_bwd_sums(dg1.ptr, db1.ptr, sums1.ptr, bound1.ptr, dy.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr,
                             mean.ptr, invstd.ptr, n, c, hw, act)
    lib.ng_bn2d_act_bwd_dx(dx1.ptr, dy.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, mean.ptr, invstd.ptr, sums1.ptr,
                           n, c, hw, act)
    return lib, dev, (mean, invstd), (pooled, dpool), (dg1, db1, sums1, bound1, dx1)


def _check_pooled_sums(shape, act):
    from nanograd_b200.nn.buffer import DeviceBuffer
    n, c, h, w = shape
    lib, dev, (mean, invstd), (pooled, dpool), (dg1, db1, sums1, bound1, dx1) = _pooled_setup(shape, act, seed=3 * c + act)
    dg2, db2, sums2, bound2 = DeviceBuffer((c,)), DeviceBuffer((c,)), DeviceBuffer((2 * c + 1,)), DeviceBuffer((1,))
    lib.ng_bn2d_act_bwd_sums_pool(dg2.ptr, db2.ptr, sums2.ptr, bound2.ptr, dpool.ptr, pooled.ptr, dev["x"].ptr, dev["gamma"].ptr,
                                  dev["beta"].ptr, mean.ptr, invstd.ptr, n, c, h, w, act)
    # same terms, another summation order (windows instead of rows): compare on the scale of sum |term|
    for name, a, b in (("dgamma", dg2, dg1), ("dbeta", db2, db1), ("sums", sums2, sums1)):
        av, bv = a.numpy().astype(np.float64), b.numpy().astype(np.float64)
        np.testing.assert_allclose(av, bv, rtol=1e-4, atol=1e-5 * max(float(np.abs(bv).max()), 1e-30), err_msg=name)
    bnd = float(bound2.numpy().view(np.float32)[0])
    dxmax = float(np.abs(dx1.numpy()).max())
    assert dxmax <= bnd * 1.001 <= 32 * dxmax, (bnd, dxmax)
    return lib, dev, mean, invstd, pooled, dpool, sums2, bound2, dx1


@pytest.mark.parametrize("act", [0, 1])
def test_bn_backward_sums_with_fused_maxpool_backward_emulated(act):
    H.use_emulated_library()
    _check_pooled_sums((2, 6, 4, 8), act)


@pytest.mark.gpu
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("shape", [(2, 6, 4, 8), (16, 64, 32, 32), (8, 128, 16, 16), (5, 192, 8, 8), (3, 64, 4, 24)],
                         ids=lambda s: "x".join(map(str, s)))
def test_bn_backward_with_fused_maxpool_backward(shape, act):
    """conv -> BatchNorm2d (+ReLU) -> MaxPool2d(2) backward without ever writing the gradient of the BatchNorm output:
    reduction (ng_bn2d_act_bwd_sums_pool) and, for 64-multiple channel counts, the staged FP16-pair copy of dx
    (ng_conv2d_stage_f16_bnbwd_pool) against the materialised chain pool backward -> sums -> dx."""
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    n, c, h, w = shape
    lib, dev, mean, invstd, pooled, dpool, sums2, bound2, dx1 = _check_pooled_sums(shape, act)
    if c % 64:
        return
    hw = h * w
    staged, chan = DeviceBuffer((n * c * hw + 16,)), DeviceBuffer((c,))
    lib.ng_conv2d_stage_f16_bnbwd_pool(staged.ptr, chan.ptr, dpool.ptr, pooled.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr,
                                       mean.ptr, invstd.ptr, sums2.ptr, bound2.ptr, n, c, h, w, act)
    got, inv_scale = _decode(staged, n, c, hw)
    dx = dx1.numpy().astype(np.float64).reshape(n, c, hw)
    dmax = np.abs(dx).max()
    assert 2.0 ** 9 <= dmax / inv_scale < 2.0 ** 15
    err = np.abs(got - dx)
    # (sums2 and the materialised chain's sums differ by a summation order: 1e-5 of the scale on top of the FP16 pairs)
    assert np.all(err <= np.maximum(np.abs(dx) * 2.0 ** -20, dmax * 2e-5)), float(err.max() / dmax)
    np.testing.assert_allclose(chan.numpy(), dx.sum(axis=(0, 2)), rtol=0, atol=1e-4 * float(np.abs(dx).sum(axis=(0, 2)).max()),
                               err_msg="channel sums of the fused dx staging")


def _check_dx_with_channel_sums(shape, act):
    """ng_bn2d_act_bwd_dx_sum: the dx of ng_bn2d_act_bwd_dx bit for bit, plus its channel sums (the bias gradient of
    the convolution below, which vanishes analytically: compared on the scale of sum |dx|)."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    lib = B.lib()
    n, c, h, w = shape
    hw = h * w
    host, dev = _setup(n, c, h, w, seed=5 * c + act)
    one, _, (_, sums, _, mean, invstd) = _one_call_and_split(lib, dev, n, c, hw, act)
    dx, chan = DeviceBuffer((n, c, hw)), DeviceBuffer((c,))
    lib.ng_bn2d_act_bwd_dx_sum(dx.ptr, chan.ptr, dev["dy"].ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr if act else 0,
                               mean.ptr, invstd.ptr, sums.ptr, n, c, hw, act)
    H.assert_exact(dx.numpy(), one[3], "dx of the pass with channel sums")
    want = one[3].astype(np.float64).sum(axis=(0, 2))
    np.testing.assert_allclose(chan.numpy(), want, rtol=0, atol=1e-5 * float(np.abs(one[3]).astype(np.float64).sum(axis=(0, 2)).max()))


@pytest.mark.parametrize("act", [0, 1])
def test_dx_pass_with_channel_sums_emulated(act):
    H.use_emulated_library()
    _check_dx_with_channel_sums((3, 8, 5, 4), act)


@pytest.mark.gpu
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("shape", [(3, 8, 5, 4), (16, 64, 32, 32), (7, 192, 9, 5)], ids=lambda s: "x".join(map(str, s)))
def test_dx_pass_with_channel_sums(shape, act):
    H.use_cuda_library()
    _check_dx_with_channel_sums(shape, act)


def test_recipe_is_immune_to_in_place_parameter_updates_emulated():
    """The optimizers update parameters in place; a BatchNorm output that is materialised only after the step (a user
    reading an intermediate activation late) must still hold the values of the step that produced it: the recipe
    carries its own snapshot of gamma / beta."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn import ops_cuda
    from nanograd_b200.nn.buffer import DeviceBuffer, LazyBuffer
    H.use_emulated_library()
    n, c, h, w = 2, 64, 2, 4
    host, dev = _setup(n, c, h, w, seed=9)

    class Ctx:
        def save_for_backward(self, *a):
            self.saved = a

    out = ops_cuda.BatchNorm2d.forward(Ctx(), dev["x"], dev["gamma"], dev["beta"], None, None, training=True, relu=True)
    assert isinstance(out, LazyBuffer) and out.pending
    eager = DeviceBuffer((n, c, h, w))
    m, i = DeviceBuffer((c,)), DeviceBuffer((c,))
    B.lib().ng_bn2d_act_fwd_train(eager.ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr, 0, 0, m.ptr, i.ptr, n, c, h * w,
                                  1e-5, 0.1, 1)
    # the backward recipe (BatchNorm's input gradient) as well
    ctx = Ctx()
    ctx.parents = ()
    out2 = ops_cuda.BatchNorm2d.forward(ctx, dev["x"], dev["gamma"], dev["beta"], None, None, training=True, relu=True)
    ctx.saved_tensors = list(ctx.saved)
    needs = ops_cuda._needs
    ops_cuda._needs = lambda c_, i_: True      # a bare ctx has no parents to ask
    try:
        dx, dg, db = ops_cuda.BatchNorm2d.backward(ctx, dev["dy"])
    finally:
        ops_cuda._needs = needs
    assert isinstance(dx, LazyBuffer) and dx.pending
    dx_eager, dg2, db2 = DeviceBuffer((n, c, h, w)), DeviceBuffer((c,)), DeviceBuffer((c,))
    B.lib().ng_bn2d_act_bwd(dx_eager.ptr, dg2.ptr, db2.ptr, dev["dy"].ptr, dev["x"].ptr, dev["gamma"].ptr, dev["beta"].ptr,
                            m.ptr, i.ptr, n, c, h * w, 1)
    dev["gamma"].copy_from_host(host["gamma"] * 3 + 1)      # "optimizer.step()"
    dev["beta"].copy_from_host(host["beta"] - 2)
    H.assert_exact(out.numpy(), eager.numpy(), "late materialisation of a BatchNorm recipe")
    H.assert_exact(dx.numpy(), dx_eager.numpy(), "late materialisation of a BatchNorm gradient recipe")
