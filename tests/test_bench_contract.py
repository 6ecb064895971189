"""CPU: the reference arm of bench.py (the one arm that runs without a GPU) prints exactly one JSON line on stdout
with the keys the driver reads; everything else goes to stderr."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("workload,metric_word", [("c1", "queries/sec"), ("c5", "pairs/sec")])
def test_reference_arm_prints_one_json_line(workload, metric_word):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and metric_word in d["metric"]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and "workload" in d["config"]
