"""bench.py's CPU arm (`--impl reference`) runs here without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--batch", "24", "--ref-sample", "8"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "nmpc_solves_per_sec" and line["unit"] == "solves/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    # N=20,Nobs=10 -> 4128 B; N=40 -> 7808 B; N=20,Nobs=50 -> 5088 B; N=10 -> 2288 B; N=80,Nobs=200 -> 19728 B
    assert bench.algorithmic_bytes(20, 10, 3) == 4128
    assert bench.algorithmic_bytes(40, 10, 3) == 7808
    assert bench.algorithmic_bytes(20, 50, 3) == 5088
    assert bench.algorithmic_bytes(10, 10, 3) == 2288
    assert bench.algorithmic_bytes(80, 200, 3) == 19728
