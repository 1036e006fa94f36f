"""bench.py contract (CPU side): the reference arm runs without a GPU and prints ONE JSON line with the keys the
driver reads; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "2"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "volumes/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_are_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_arm_is_like_for_like_with_the_gpu_arm():
    """same config keys / batch as the GPU arm, the driver's --steps and --warmup honoured."""
    sys.path.insert(0, ROOT)
    import bench
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "2",
                        "--batch", "2", "--gpus", "4"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["steps"] == 3 and d["warmup"] == 2 and d["n_gpus"] == 4
    assert d["config"] == bench.workload_config(bench.WORKLOADS["conf5_infer"], 2, 4, 1)
    assert set(d["config"]) == {"workload", "batch_per_gpu", "global_batch", "vis", "l2", "parallelism"}


def test_timed_regions_do_not_call_the_oracle():
    """bench.py may use oracle/ for synthetic inputs, seeded weights and the CPU arm only - never inside a step."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def run_leg"):src.index("def ensemble_small_batch_probe")]
    step_fns = body[body.index("def pos_weight"):body.index("ms_dev, launches")]
    assert "O." not in step_fns and "oracle" not in step_fns
