"""bench.py keeps the driver's contract: the CPU reference arm runs here (no GPU) and prints ONE JSON line with the agreed keys; the
committed line of the GPU arm (profiles/r2_bench_final.json, written by scripts/profile_pass.sh on a B200) carries the roofline,
cpu_baseline, e2e and clock objects the contract asks for."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "patches/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-9
    e = d["e2e"]
    assert abs(e["value"] - d["value"]) < 1e-9 and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_committed_gpu_line_has_the_contract_objects():
    p = ROOT / "profiles" / "r2_bench_final.json"
    d = json.loads(p.read_text().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["dtype"] == "bf16" and d["scaling"] == "weak" and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "TFLOP/s"
    assert r["traffic"] is None or r["traffic"] > 0
    rl = d["roofline_loss"]
    assert rl["bound"] == "hbm" and abs(rl["frac"] - rl["achieved"] / rl["peak"]) < 1e-9 and rl["frac"] >= 0.60      # the north star's loss target
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
