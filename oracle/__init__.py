"""TEST INFRASTRUCTURE — CPU oracle for the nanograd hot path.  NOT part of the product.

A NumPy restatement of the algorithms of PABannier/nanograd's CPU path
(``nanograd/nn/ops_cpu.py``, ``nanograd/tensor.py`` composites, ``nanograd/nn/module.py`` layers,
``nanograd/optim/optimizer.py``), each function citing the reference lines it follows.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` — as the checker or the timed CPU baseline only.  The
``nanograd_b200`` package never imports it; the device path has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_pin.py`` checks every function here against
(a) the committed golden vectors in ``tests/golden/*.npz``, which were produced by importing the
real reference from /root/reference (generator: ``tests/golden/make_golden.py``), and (b) the
real reference itself whenever /root/reference is present (the build container).  The
reference's own tests compare against PyTorch-CPU only and ship no golden files.
"""


def register_host():
    """Plug the oracle into this repo's host mirror as its ``Device.CPU`` side: the primitive
    namespace, the BatchNorm compositions and the optimizer steps.  Tests / smoke / bench
    baselines only."""
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import register_ops
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    from oracle import cpu_namespace, host_layers
    register_ops(cpu_namespace, device=Device.CPU)
    nn.set_host_forward("BatchNorm1d", host_layers.batchnorm1d_forward)
    nn.set_host_forward("BatchNorm2d", host_layers.batchnorm2d_forward)
    optim.set_host_step("sgd", host_layers.sgd_step)
    optim.set_host_step("adam", host_layers.adam_step)
    optim.set_host_step("adamw", host_layers.adamw_step)
