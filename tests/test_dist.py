"""Data-parallel path: sharding arithmetic; at world size 2 on CPU (gloo) the host-side
gradient math and the full device code path on the emulated kernels (gradient arena, chunked allreduce,
SyncBN, fused optimizer); and — on a box with >= 2 B200s (`gpurun --gpus 2`) — the same over NCCL against the
single-process oracle on the global batch."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(mode, nproc=2, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER, mode]
    env = dict(os.environ, OMP_NUM_THREADS="2", OPENBLAS_NUM_THREADS="2")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_shard_bounds_partition_the_batch():
    from nanograd_b200 import dist
    for batch, world in [(4096, 8), (512, 1), (10, 4), (7, 2), (3, 8)]:
        spans = [dist.shard_bounds(batch, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_shard_loss_weight_restores_the_global_batch_gradient():
    """Unequal shards: mean-loss gradients averaged with 1/world are not the global-batch gradient unless each
    rank's loss carries n_r * world / N (1.0 for the equal shards of every BASELINE configuration)."""
    import numpy as np
    from nanograd_b200 import dist
    assert dist.shard_loss_weight(4096, 8, 3) == 1.0 and dist.shard_loss_weight(512, 1, 0) == 1.0
    rng = np.random.RandomState(0)
    for batch, world in [(10, 4), (7, 2), (33, 8)]:
        X, y, w = rng.randn(batch, 5), rng.randn(batch), rng.randn(5)
        grad = lambda lo, hi: 2.0 * X[lo:hi].T @ (X[lo:hi] @ w - y[lo:hi]) / (hi - lo)   # d/dw mean((Xw - y)^2)
        want = grad(0, batch)
        spans = [dist.shard_bounds(batch, world, r) for r in range(world)]
        plain = sum(grad(lo, hi) for lo, hi in spans) / world
        weighted = sum(dist.shard_loss_weight(batch, world, r) * grad(lo, hi) for r, (lo, hi) in enumerate(spans)) / world
        assert np.allclose(weighted, want, rtol=1e-12, atol=1e-12)
        assert not np.allclose(plain, want, rtol=1e-6)


def test_world2_gloo_control_plane_and_dp_gradient_math():
    res = _torchrun("cpu")
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_OK 0" in res.stdout and "DIST_OK 1" in res.stdout, res.stdout[-2000:]


def test_world2_emulated_dp_path_matches_global_batch_oracle():
    """The whole data-parallel code path (gradient arena with in-place slots and chunked allreduce,
    synchronised BatchNorm, fused optimizer) on the host emulation of the kernels over gloo."""
    res = _torchrun("emu", timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_OK 0" in res.stdout and "DIST_OK 1" in res.stdout, res.stdout[-2000:]


@pytest.mark.gpu
def test_world2_nccl_step_matches_global_batch_oracle():
    try:
        n = len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
    except Exception:
        n = 0
    if n < 2:
        pytest.skip("needs 2 GPUs on the box (run with gpurun --gpus 2)")
    res = _torchrun("gpu")
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_OK 0" in res.stdout and "DIST_OK 1" in res.stdout, res.stdout[-2000:]
