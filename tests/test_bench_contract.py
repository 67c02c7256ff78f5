"""bench.py's JSON contract on the one arm that runs without a GPU: `--impl reference` falls back to the CPU oracle port when no
CUDA device is present (the product arm refuses to run without one: the rasterizer has no CPU path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line_without_a_gpu():
    r = _run("--impl", "reference", "--config", "plumbing-100k", "--steps", "1", "--warmup", "0", "--cpu-tiles", "4")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "views/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
    assert line["e2e"] == {"value": cb["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)
