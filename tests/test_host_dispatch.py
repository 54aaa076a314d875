"""Host logic of the tensor-core step path, without a GPU.

The emulated library (tests/emu) has no tensor-core kernels, so on the CPU-only run the host code that drives the
FP16x3 engine — operand staging, BatchNorm recipes (LazyBuffer), pooled recipes — would never execute.  Here the
product's Python path runs BASELINE config 2 at its full batch against a *null* build of the C ABI
(tools/host_dispatch_prof.py: every entry point returns at once, ng_conv2d_tc_plan answers like the sm_100a
library), in a subprocess because a process binds one library.  Checked: the step dispatches without a Python error
through the real ctypes prototypes, it makes exactly the library calls DESIGN.md §3.4/§3.5 describe (nothing is
materialised in a conv -> BatchNorm -> ReLU -> conv / max-pool chain), and it leaves no device allocation behind.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    env = dict(os.environ)
    env.pop("NANOGRAD_B200_MATH", None)
    env.pop("NANOGRAD_TF32", None)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "host_dispatch_prof.py"), "--json", "--steps", "3"] + list(flags),
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout + res.stderr
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_default_step_takes_the_fused_tensor_core_path():
    r = _run("--model", "vgg")
    c = r["calls"]
    assert r["batch"] == 512 and r["math"] == "fp32"
    # six convolutions forward and six wgrads; five dgrads: the image does not ask for a gradient (ops_cuda._needs)
    assert c["ng_conv2d_fprop"] == 6 and c["ng_conv2d_wgrad"] == 6 and c["ng_conv2d_dgrad"] == 5
    # BatchNorm: statistics only, forward; sums only, backward (three of them with the max-pool backward fused in);
    # the first layer's dx is the one tensor that is materialised (with the conv bias gradient in the same pass)
    assert c["ng_bn2d_act_fwd_stats"] == 6
    assert c["ng_bn2d_act_bwd_sums"] + c["ng_bn2d_act_bwd_sums_pool"] == 6 and c["ng_bn2d_act_bwd_sums_pool"] == 3
    assert c["ng_bn2d_act_bwd_dx_sum"] == 1
    for never in ("ng_bn2d_act_apply", "ng_bn2d_act_bwd_dx", "ng_bn2d_act_fwd_train", "ng_bn2d_act_bwd",
                  "ng_maxpool2d_fwd", "ng_maxpool2d_bwd", "ng_maxpool2d_bwd_bn", "ng_conv2d_stage_nhwc"):
        assert never not in c, never
    # operand staging: x of the five tensor-core layers (three from BatchNorm recipes, two from pooled tensors),
    # dy of the same five (recipes of BatchNorm's backward, three of them through a max-pool)
    assert c["ng_conv2d_stage_f16_bn"] == 3 and c["ng_conv2d_stage_f16"] == 2
    assert c["ng_conv2d_stage_f16_bnbwd"] + c["ng_conv2d_stage_f16_bnbwd_pool"] == 5
    assert c["ng_conv2d_stage_filters_f16"] == 5
    assert c["ng_maxpool2d_fwd_bn"] == 3
    assert c["ng_gemm"] == 6 and c["ng_multi_adam"] == 1
    assert c["ng_logsoftmax_fwd"] == c["ng_logsoftmax_bwd"] == c["ng_nll_fwd"] == c["ng_nll_bwd"] == 1
    # a step returns every allocation it made
    assert r["allocs_per_step"] > 0
    assert r["live_growth_one_step"] == 0 and r["live_growth_timed_steps"] == 0


@pytest.mark.parametrize("mode", ["simt", "tf32", "fp32_tf32x3", "f16x3"])
def test_other_math_modes_dispatch_and_free(mode):
    r = _run("--model", "vgg", "--math", mode)
    c = r["calls"]
    assert c["ng_conv2d_fprop"] == 6 and c["ng_conv2d_wgrad"] == 6 and c["ng_conv2d_dgrad"] == 5
    if mode == "simt":
        assert not [k for k in c if k.startswith("ng_conv2d_stage")]
        assert c["ng_bn2d_act_fwd_train"] == 6 and c["ng_bn2d_act_bwd"] == 6
    if mode in ("tf32", "fp32_tf32x3"):
        assert c["ng_conv2d_stage_nhwc"] == 5 and c["ng_conv2d_stage_nhwc_sum"] == 5
    assert r["live_growth_one_step"] == 0 and r["live_growth_timed_steps"] == 0


def test_readme_cnn_step_dispatch():
    r = _run("--model", "mnist")
    c = r["calls"]
    assert c["ng_conv2d_fprop"] == 2 and c["ng_conv2d_wgrad"] == 2 and c["ng_conv2d_dgrad"] == 1
    assert c["ng_maxpool2d_fwd"] == 2 and c["ng_maxpool2d_bwd"] == 2 and c["ng_multi_sgd"] == 1
    assert r["library_calls_per_step"] <= 30
    assert r["live_growth_one_step"] == 0 and r["live_growth_timed_steps"] == 0


@pytest.mark.parametrize("world,sync_bn", [(2, False), (8, False), (8, True)])
def test_data_parallel_step_dispatch(world, sync_bn):
    """Rank 0 of a data-parallel group over the null library (its collectives are no-ops): gradients go straight
    into the flat arena (no pack copy), the arena is all-reduced in chunks while the backward runs and the
    optimizer waits for the tail once; SyncBN keeps the one-call BatchNorm kernels."""
    r = _run("--model", "vgg", "--world", str(world), *(["--sync-bn"] if sync_bn else []))
    c = r["calls"]
    assert r["world"] == world
    assert c["ng_nccl_allreduce_begin"] >= 2, "the gradient arena must be reduced in several chunks during backward"
    assert c["ng_nccl_allreduce_end"] == 1 and c["ng_multi_adam"] == 1
    assert "ng_multi_pack" not in c and "ng_multi_unpack" not in c and "ng_nccl_allreduce_sum" not in c
    assert c["ng_conv2d_fprop"] == 6 and c["ng_conv2d_wgrad"] == 6 and c["ng_conv2d_dgrad"] == 5
    if sync_bn:
        assert c["ng_bn2d_act_fwd_train"] == 6 and c["ng_bn2d_act_bwd"] == 6 and "ng_bn2d_act_fwd_stats" not in c
    else:
        assert c["ng_bn2d_act_fwd_stats"] == 6 and c["ng_conv2d_stage_f16_bn"] == 3
    assert r["live_growth_one_step"] == 0 and r["live_growth_timed_steps"] == 0


def test_reference_classes_reach_the_same_fused_path():
    """The unmodified reference's Tensor / Module / Optimizer, bound with nanograd_b200.integrate.install(), issue the
    same library calls per step as this package's host mirror (the backward seed apart: the reference uploads
    np.ones, tensor.py:226) — the drop-in reaches the fused tensor-core path without edits to user code."""
    import helpers
    if helpers.load_reference() is None:
        pytest.skip("the reference is not present (oracle/_ref is installed by __graft_entry__.build())")
    ours = _run("--model", "vgg")["calls"]
    theirs = _run("--model", "vgg", "--reference")
    assert theirs["reference_classes"] is True
    c = dict(theirs["calls"])
    assert c.pop("ng_h2d") == 1 and "ng_fill" not in c
    ours.pop("ng_fill")
    assert c == ours
    assert theirs["live_growth_one_step"] == 0 and theirs["live_growth_timed_steps"] == 0
