"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm) prints ONE JSON line with the contract's keys.
Runs here on the CPU with a shortened sample (GB_BENCH_REF_SECONDS); the timing itself is not asserted."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, GB_BENCH_REF_SECONDS="0.5")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"], env=env,
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # stdout carries the JSON line and nothing else (Grid's banner goes to stderr)
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFlop/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFlop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
