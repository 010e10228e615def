"""bench.py's reference arm (the CPU port of the train step) prints the contract's JSON line; runs without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-batch", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "train pairs/s (chairs_uflow)" and line["unit"] == "pairs/s" and line["higher_is_better"]
    assert line["steps"] == 1 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["value"] > 0 and abs(line["value"] - 1e3 / line["ms_per_step"]) < 1e-6 * line["value"] + 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "384x512" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]
