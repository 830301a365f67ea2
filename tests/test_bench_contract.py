"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm prints exactly
ONE JSON line on stdout (library banners and progress go to stderr) with the keys the driver
reads, non-zero ranks of a multi-rank reference launch exit 0 without work, and the CUDA arm
refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=timeout)


@pytest.mark.needs_pyhf
def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--no-toys"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["higher_is_better"] is True and d["unit"] == "evals/s" and d["value"] > 0
    assert d["metric"].startswith("batched logpdf+grad evals/s")
    assert d["config"]["workload"].startswith("C4 synthetic") and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["fd_equivalent_value"] == pytest.approx(d["value"] / 301)   # P + 1 forward differences
    # both arms print the SAME config object (one function builds it; the driver compares the two)
    import re

    src = open(os.path.join(ROOT, "bench.py")).read()
    assert len(re.findall(r'"config": bench_config\(', src)) == 2
    assert set(d["config"]) == {"workload", "batch_per_gpu", "parallelism", "l2"}
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0"], timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr
