"""CPU: the output contract of bench.py that the driver depends on -- exactly one JSON line on
stdout with the required keys -- exercised through the reference arm (the GPU arm needs a B200;
its line is checked by the same key list in the GPU suite)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"]


def _one_json_line(stdout):
    lines = [ln for ln in stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, BGP_BENCH_REF_BUDGET_S="4")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = _one_json_line(out.stdout)
    for k in REQUIRED:
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "LML evals/s" and line["value"] > 0
    assert line["config"]["workload"].startswith("C3") and "model" not in line["config"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["vs_baseline"] is None and line["higher_is_better"] is True


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", BGP_BENCH_REF_BUDGET_S="4")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120,
                         cwd=REPO)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = _one_json_line(out.stdout)
    for k in REQUIRED + ["roofline", "gpu_launches", "clocks"]:
        assert k in line, k
    assert line["gpu_launches"] > 0 and 0 < line["roofline"]["frac"] < 1
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["roofline"]["bound"] == "tensor" and line["roofline"]["unit"] == "TFLOP/s"
