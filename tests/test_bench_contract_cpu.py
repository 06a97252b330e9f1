"""bench.py contract pieces that run without a GPU: the reference arm (CPU oracle) prints exactly one JSON line
with the keys the driver reads, and the main arm refuses to run without a CUDA device instead of falling back."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "column-timesteps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("column-timesteps/sec") and d["dtype"] == "f64" and d["scaling"] in ("weak", "strong")
    assert d["value"] > 1e3 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]
    # same `config` as the repo arm prints for this launch (the sample lives in cpu_baseline): the driver compares them
    sys.path.insert(0, str(ROOT))
    import bench
    assert d["config"] == bench.workload_config(1, scaling="weak", per_gpu=bench.TOTAL_COLUMNS)
    assert "sample" not in d["config"] and "libm" in cb["sample"] and cb["det_build_value"] > 1e3


def test_main_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
    assert not r.stdout.strip()
