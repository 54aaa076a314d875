"""Worker for tests/test_dist.py, launched by torchrun with 2 ranks.

mode "cpu": gloo control plane only.  Each rank runs the README MNIST CNN (no BatchNorm, so the
            data-parallel step is exactly the global-batch step) on ITS shard with the NumPy oracle,
            gradients are summed with dist.all_reduce_host and averaged; every rank checks the result
            against the global-batch gradient it computes locally.
mode "gpu": the real path: each rank owns one B200, shard on Device.GPU, optimizer.step() packs the
            gradients, runs ONE NCCL allreduce and the fused multi-tensor update; the post-step weights
            are checked against the single-process NumPy oracle on the global batch.
Prints "DIST_OK <rank>" on success.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from nanograd_b200 import dist  # noqa: E402
from nanograd_b200.device import Device  # noqa: E402
from nanograd_b200.tensor import Tensor  # noqa: E402
import nanograd_b200.nn.module as nn  # noqa: E402
import nanograd_b200.optim.optimizer as optim  # noqa: E402
import oracle  # noqa: E402

GLOBAL_BATCH = 32


def make(device, opt_name):
    np.random.seed(0)  # identical replicas: host-side NumPy init under one seed
    model = H.build_mnist_cnn(nn)
    if device == Device.GPU:
        model.gpu()
    opt = optim.SGD(model.parameters(), lr=0.05, momentum=0.9) if opt_name == "sgdm" else optim.Adam(model.parameters(), lr=1e-3)
    return model, opt


def step(model, opt, X, Y, device):
    out = model(Tensor(X, device=device))
    opt.zero_grad()
    loss = nn.NLLLoss()(out, Tensor(Y, device=device))
    loss.backward()
    return loss


def main():
    mode = sys.argv[1]
    oracle.register_host()  # the oracle: checker, never the measured path
    rng = np.random.RandomState(7)
    X = rng.rand(GLOBAL_BATCH, 784).astype(np.float32)
    Y = rng.randint(0, 10, GLOBAL_BATCH).astype(np.float32)
    dist.init(use_device=(mode == "gpu"))
    r, w = dist.rank(), dist.world_size()
    assert w == 2, w
    lo, hi = dist.shard_bounds(GLOBAL_BATCH)
    assert (hi - lo) == GLOBAL_BATCH // w
    dist.barrier()
    assert dist.all_reduce_max(float(r)) == float(w - 1)
    assert dist.all_reduce_sum(1.0) == float(w)

    # global-batch reference on the oracle
    ref_model, ref_opt = make(Device.CPU, "sgdm")
    step(ref_model, ref_opt, X, Y, Device.CPU)
    ref_grads = [np.array(p.grad.data, dtype=np.float64) for p in ref_model.parameters()]
    ref_opt.step()
    ref_params = [np.array(p.data, dtype=np.float64) for p in ref_model.parameters()]

    if mode == "cpu":
        model, opt = make(Device.CPU, "sgdm")
        step(model, opt, X[lo:hi], Y[lo:hi], Device.CPU)
        for p, g_ref in zip(model.parameters(), ref_grads):
            g = dist.all_reduce_host(np.array(p.grad.data, dtype=np.float64)) / w
            H.assert_close(g, g_ref, 1e-6, "dp gradient")
    else:
        model, opt = make(Device.GPU, "sgdm")
        for _ in range(1):
            step(model, opt, X[lo:hi], Y[lo:hi], Device.GPU)
            opt.step()  # pack -> ncclAllReduce(sum) -> fused update with grad_scale = 1/world
        for p, p_ref in zip(model.parameters(), ref_params):
            H.assert_close(p.numpy(), p_ref, H.RTOL_FP32, "dp post-step parameter")
        # replicas must stay bit-identical across ranks
        for p in model.parameters():
            a = p.numpy().astype(np.float64)
            s = dist.all_reduce_host(a.copy())
            assert np.array_equal(s, a * w), "replicas diverged"
    dist.barrier()
    print(f"DIST_OK {r}", flush=True)
    dist.shutdown()


if __name__ == "__main__":
    main()
