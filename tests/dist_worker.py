"""Worker for tests/test_dist.py, launched with 2 ranks (torchrun).

mode "cpu": gloo control plane only.  Each rank runs the README MNIST CNN (no BatchNorm, so the
            data-parallel step is exactly the global-batch step) on ITS shard with the NumPy oracle,
            gradients are summed with dist.all_reduce_host and averaged; every rank checks the result
            against the global-batch gradient it computes locally.
mode "emu": the real data-parallel code path on the host emulation of the kernels, collectives carried
            by gloo: gradient arena with in-place slots and chunked allreduce, synchronised BatchNorm,
            fused optimizer with grad_scale = 1/world — for the MNIST CNN (SGD momentum) and a
            BatchNorm VGG (Adam / SGD), 3 steps each, against the single-process NumPy oracle on the
            GLOBAL batch.
mode "gpu": the same checks on two B200s over NCCL (file rendezvous, no torch.distributed), in both
            FP32-accurate engines, plus bit-identical replicas.
Prints "DIST_OK <rank>" on success.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
import oracle  # noqa: E402
from nanograd_b200 import dist  # noqa: E402
from nanograd_b200.device import Device  # noqa: E402
from nanograd_b200.tensor import Tensor  # noqa: E402
import nanograd_b200.nn.module as nn  # noqa: E402
import nanograd_b200.optim.optimizer as optim  # noqa: E402

GLOBAL_BATCH = 32
STEPS = 3   # 2 on the host emulation (set in main)


def build(kind):
    np.random.seed(0)  # identical replicas: host-side NumPy init under one seed
    if kind == "mnist":
        return H.build_mnist_cnn(nn)
    body = H.build_vgg(nn, widths=(8, 16), image=8, hidden=32)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.body = body

        def forward(self, x):
            return self.body(x).log_softmax()
    return Net()


def make_opt(model, name):
    params = list(model.parameters())
    return optim.SGD(params, lr=0.05, momentum=0.9) if name == "sgdm" else optim.Adam(params, lr=1e-3)


def data(kind):
    rng = np.random.RandomState(7)
    if kind == "mnist":
        X = rng.rand(STEPS, GLOBAL_BATCH, 784).astype(np.float32)
    else:
        X = rng.randn(STEPS, GLOBAL_BATCH, 3, 8, 8).astype(np.float32)
    Y = rng.randint(0, 10, (STEPS, GLOBAL_BATCH)).astype(np.float32)
    return X, Y


def run(model, opt, X, Y, device, lo, hi):
    losses = []
    for s in range(STEPS):
        out = model(Tensor(X[s, lo:hi], device=device))
        opt.zero_grad()
        loss = nn.NLLLoss()(out, Tensor(Y[s, lo:hi], device=device))
        loss.backward()
        opt.step()
        losses.append(float(np.asarray(loss.numpy()).reshape(-1)[0]))
    return losses


def host(p):
    return np.asarray(p.numpy(), dtype=np.float64)


def check_dp(kind, opt_name, rtol, what):
    """Data-parallel run on Device.GPU (sharded batch) against the oracle on the global batch."""
    X, Y = data(kind)
    ref_model = build(kind)
    ref_opt = make_opt(ref_model, opt_name)
    ref_losses = run(ref_model, ref_opt, X, Y, Device.CPU, 0, GLOBAL_BATCH)
    lo, hi = dist.shard_bounds(GLOBAL_BATCH)
    model = build(kind).gpu()
    opt = make_opt(model, opt_name)
    losses = run(model, opt, X, Y, Device.GPU, lo, hi)
    # the shard losses average to the global loss (NLL is a mean: module.py:686)
    mean_losses = [dist.all_reduce_sum(l) / dist.world_size() for l in losses]
    np.testing.assert_allclose(mean_losses, ref_losses, rtol=max(rtol, 1e-5), err_msg=f"{what}: loss history")
    ref_grad_max = max(float(np.abs(host(p.grad)).max()) for p in ref_model.parameters())
    for i, (p, q) in enumerate(zip(model.parameters(), ref_model.parameters())):
        if float(np.abs(host(q.grad)).max()) < 1e-6 * ref_grad_max:
            # analytically-zero gradient (conv bias before BatchNorm): the parameter only moves by rounding
            # noise (SGD) or by noise amplified to ~lr per step (Adam) on every implementation
            assert float(np.abs(host(p) - host(q)).max()) <= (1e-6 if opt_name != "adam" else 1e-2), what
            continue
        H.assert_close(host(p), host(q), rtol, f"{what}: parameter {i} after {STEPS} steps")
    if kind == "vgg":
        bns = [(m, r) for m, r in zip(model.body.layers, ref_model.body.layers) if type(m).__name__ == "BatchNorm2d"]
        assert bns
        for m, r in bns:
            H.assert_close(host(m.running_mean), host(r.running_mean), rtol, f"{what}: running mean")
            H.assert_close(host(m.running_var), host(r.running_var), rtol, f"{what}: running var")
    # replicas stay bit-identical
    chk = dist.replica_checksum(list(model.parameters()))
    assert chk["identical_across_ranks"], f"{what}: replicas diverged"
    # the gradients were produced in place in the arena from step 2 on
    bucket = opt._bucket
    assert bucket is not None and bucket.total >= sum(bucket.numels)
    in_place = sum(1 for i, p in enumerate(model.parameters()) if p.grad.data.ptr == bucket.slot_ptr(i))
    assert in_place >= len(bucket.numels) - 1, f"{what}: only {in_place}/{len(bucket.numels)} gradients were produced in place"


def main():
    mode = sys.argv[1]
    oracle.register_host()  # the oracle: checker, never the measured path
    if mode == "cpu":
        dist.init(use_device=False)
    elif mode == "emu":
        global STEPS, GLOBAL_BATCH
        STEPS, GLOBAL_BATCH = 2, 8
        H.use_emulated_library()
        dist.init(use_device="emu", sync_bn=True)
    else:
        H.use_cuda_library()
        dist.init(use_device=True, sync_bn=True)
    r, w = dist.rank(), dist.world_size()
    assert w == 2, w
    lo, hi = dist.shard_bounds(GLOBAL_BATCH)
    assert (hi - lo) == GLOBAL_BATCH // w
    if mode == "emu":
        assert "torch" in sys.modules          # gloo carries the emulated collectives
    if mode == "gpu":
        assert "torch" not in sys.modules, "the GPU data-parallel path must not import torch"
    dist.barrier()
    assert dist.all_reduce_max(float(r)) == float(w - 1)
    assert dist.all_reduce_sum(1.0) == float(w)

    if mode == "cpu":
        X, Y = data("mnist")
        ref_model = build("mnist")
        out = ref_model(Tensor(X[0]))
        nn.NLLLoss()(out, Tensor(Y[0])).backward()
        ref_grads = [np.array(p.grad.data, dtype=np.float64) for p in ref_model.parameters()]
        model = build("mnist")
        nn.NLLLoss()(model(Tensor(X[0, lo:hi])), Tensor(Y[0, lo:hi])).backward()
        for p, g_ref in zip(model.parameters(), ref_grads):
            g = dist.all_reduce_host(np.array(p.grad.data, dtype=np.float64)) / w
            H.assert_close(g, g_ref, 1e-6, "dp gradient")
    else:
        from nanograd_b200 import backend
        modes = ["fp32"] if mode == "emu" else ["simt", "tf32x3", "f16x3"]
        for math in modes:
            backend.set_math_mode(math)
            check_dp("mnist", "sgdm", H.RTOL_FP32, f"[{math}] MNIST CNN, SGD momentum")
            check_dp("vgg", "sgdm", H.RTOL_FP32, f"[{math}] BatchNorm VGG (SyncBN), SGD momentum")
            if mode != "emu":
                # Adam amplifies rounding differences step after step (it divides by sqrt(v) ~ |g|): measured
                # 1.3e-3 on a running mean after a mere change of summation order in a wgrad kernel
                check_dp("vgg", "adam", 3e-3, f"[{math}] BatchNorm VGG (SyncBN), Adam")
        backend.set_math_mode("fp32")
    dist.barrier()
    print(f"DIST_OK {r}", flush=True)
    dist.shutdown()


if __name__ == "__main__":
    main()
