"""bench.py's host logic and output contract, without a GPU: the device arm runs end to end over the null build of
the C ABI (tools/host_dispatch_prof.py; timings are meaningless there — event brackets return constants — but every
code path that builds the JSON line executes), single-process and as rank 0 of an 8-rank strong-scaling SyncBN run.
The numbers of record come from the B200 (profiles/r2_final_bench.json); this guards the plumbing around them."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                 "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline")


def _bench(*flags, env=None):
    import host_dispatch_prof
    lib = host_dispatch_prof.build_null_library()
    e = dict(os.environ, NANOGRAD_B200_LIB=lib)
    for k in ("NANOGRAD_B200_MATH", "NANOGRAD_TF32", "WORLD_SIZE", "RANK", "LOCAL_RANK"):
        e.pop(k, None)
    e.update(env or {})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3"] + list(flags),
                         capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly one JSON line, got %d" % len(lines)
    return json.loads(lines[0])


def test_single_gpu_line_carries_the_contract_and_the_extras():
    d = _bench()
    for k in CONTRACT_KEYS:
        assert k in d, k
    assert d["metric"] == "train_images_per_sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 512 * 3 * 32 * 32 * 4 + 512 * 4 and d["e2e"]["d2h_bytes_per_step"] == 4
    assert d["gpu_launches"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    for k in ("simt_mode", "tf32_mode", "conv2d_sweep", "mnist_cnn", "attention", "cpu_baseline"):
        assert k in d and "error" not in d[k], (k, d.get(k))
    cb = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(cb) and cb["kind"] in ("reference", "port")


def test_rank0_of_an_eight_rank_strong_scaling_syncbn_run():
    d = _bench("--gpus", "8", "--scaling", "strong", "--sync-bn", "--no-tf32",
               env={"WORLD_SIZE": "8", "RANK": "0", "LOCAL_RANK": "0", "MASTER_PORT": "1",
                    "NANOGRAD_B200_RDZV_DIR": os.environ.get("TMPDIR", "/tmp")})
    assert d["n_gpus"] == 8 and d["scaling"] == "strong"
    assert d["config"]["global_batch"] == 4096 and d["config"]["per_gpu_batch"] == 512 and d["config"]["parallelism"] == "dp8"
    assert d["replica_checksum"]["identical_across_ranks"] is True
    assert "conv2d_sweep" not in d            # side measurements are single-GPU
