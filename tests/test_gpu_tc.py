"""-m gpu: tcgen05/TMA conv and wgrad kernels vs the CUDA-core kernels, run in a subprocess with a
timeout (a mis-programmed mbarrier pipeline hangs instead of failing)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
TOL = 5e-3   # bf16 outputs: both kernels accumulate in fp32, results differ by bf16 rounding of the sum only


@pytest.fixture(scope="module")
def report():
    out = ROOT / "gpurun_out" / "tc_probe.json"
    out.parent.mkdir(exist_ok=True)
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "tc_probe.py"), str(out)], capture_output=True, text=True, timeout=600)
    (ROOT / "gpurun_out" / "tc_probe.log").write_text(r.stdout + "\n--- stderr ---\n" + r.stderr)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(out.read_text())


@pytest.mark.parametrize("variant", ["v2_auto", "v1_mt1", "v2_res_mt2", "v2_nores_mt1", "v2_nores_mt2", "v2_nores_mt4", "v2_no_ns3", "v2_no_ns3_nores",
                                     "v2_ns3", "v2_ns3_nores_one_cta", "v2_ns3_nores_two_cta", "v2_ns3_nores_mt2",
                                     "v2_ew16", "v2_ew16_ns3", "v2_ew16_nores_mt4"])
def test_conv_tc(report, variant):
    """v2_auto is what the library picks (persistent CTAs, resident weights when they fit, fused BN statistics, column taps
    stacked along N for narrow N tiles); the other variants pin the streaming / multi-window / non-persistent / one-UMMA-per-tap
    code paths.  Every variant is compared BOTH with the CUDA-core kernel and with the torch-CPU shadow of the op
    (tests/shadow_ops.py), so the tcgen05 path is not only checked against this library's own kernels."""
    bad = {k: v[variant] for k, v in report["conv"].items() if not (isinstance(v[variant], float) and v[variant] < TOL)}
    assert not bad, bad


def test_wgrad_tc(report):
    bad = {k: v for k, v in report["wgrad"].items()
           if any(key.startswith("error") for key in v) or any(val > TOL for key, val in v.items() if not key.startswith("error"))}
    assert not bad, bad
