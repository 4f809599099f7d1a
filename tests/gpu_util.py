"""Helpers for the -m gpu tests: mirror GPU views on the CPU and compare CUDA ops with the shadow ops."""
import torch

from kurosiwo_b200.lib import View


def mirror(v: View) -> View:
    return View(v.base.detach().cpu().clone(), v.offset, v.N, v.H, v.W, v.C, v.sn, v.sh, v.sw)


def rand_view(N, H, W, C, dtype, device, ctot=None, c0=0, scale=1.0, gen=None):
    """A view of C channels starting at channel c0 of a [N,H,W,ctot] buffer filled with N(0,scale)."""
    ctot = ctot or C
    base = (torch.randn(N * H * W * ctot, generator=gen, device="cpu") * scale).to(dtype).to(device)
    full = View(base, 0, N, H, W, ctot, H * W * ctot, W * ctot, ctot)
    return full.ch(c0, C), full


def bf16_drift(case: str) -> dict:
    """{parameter name: rel-L2 distance between the UNMODIFIED reference's gradient under torch.autocast(bfloat16) and its fp32
    gradient} for one of the cases of tests/golden/bf16_drift.npz (oracle/make_golden.py --bf16-drift-only).  The bf16 tests
    bound the CUDA path's gradient error by DRIFT_FACTOR x this drift + DRIFT_FLOOR: the bar is the reference's own bf16 behaviour."""
    import numpy as np
    from pathlib import Path
    d = np.load(Path(__file__).parent / "golden" / "bf16_drift.npz")
    return dict(zip([str(s) for s in d[f"{case}.names"]], [float(v) for v in d[f"{case}.drift"]]))


DRIFT_FACTOR, DRIFT_FLOOR = 1.5, 0.02


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
