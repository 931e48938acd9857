"""The reference arm of bench.py (`--impl reference`) runs without a GPU: its JSON line must carry the keys the
driver reads (same metric / unit / config as the GPU arm, `impl`, `cpu_baseline`, an `e2e` object with zero copies)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, FRS_BENCH_ROWS="100000", **extra_env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *args],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return lines


def test_reference_arm_prints_one_contract_line():
    lines = _run({})
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    assert j["metric"].startswith("exact top-15 cosine search QPS") and j["unit"] == "queries/s"
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["steps"] == 1 and j["warmup"] == 1
    assert j["config"]["workload"].startswith("100000 x 384") and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "rows per step" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    # under torchrun only rank 0 runs the CPU arm; the other ranks print nothing and exit 0
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == []
