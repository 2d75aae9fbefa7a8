"""bench.py's reference arm (`--impl reference`) runs on host cores only, so its JSON contract can be checked here: one line,
the keys the driver reads, the CPU restatement as the thing measured (kind "port"), zero transfer bytes, and ranks other than 0
leaving quietly.  (The GPU arm's line is checked on the GPU box by the driver and by tools/gpu_r02_final.sh.)"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *args],
                          capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)


def test_reference_arm_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert d["impl"] == "reference" and d["metric"] in base["metric"] and d["unit"] == "Gb/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["dtype"] == "f64" and d["config"]["workload"] == "bg1_z384_r13_it8_b4096" and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "codewords" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
