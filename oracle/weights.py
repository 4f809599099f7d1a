"""Deterministic (numpy PCG64) SNUNet_ECAM state dicts shared by the golden generator and the tests.

Key names / shapes follow the reference constructor (models/snunet.py:65-108).
Values are NOT the reference's kaiming init: BN affine terms are randomised too so that tests
exercise gamma/beta paths.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

DEC_ORDER = [(0, 1), (1, 1), (0, 2), (2, 1), (1, 2), (0, 3), (3, 1), (2, 2), (1, 3), (0, 4)]


def block_names():
    names = []
    for l in range(5):
        names.append((f"conv{l}_0", l, 0))
        if l >= 1:
            names.append((f"Up{l}_0", l, None))
    for j in range(1, 5):
        for l in range(0, 5 - j):
            names.append((f"conv{l}_{j}", l, j))
            if l >= 1 and (l - 1, j + 1) in DEC_ORDER:
                names.append((f"Up{l}_{j}", l, None))
    return names


def make_state(seed: int, in_ch: int = 2, out_ch: int = 3, base: int = 32) -> "OrderedDict[str, np.ndarray]":
    rng = np.random.Generator(np.random.PCG64(seed))
    f = [base * (1 << l) for l in range(5)]
    sd = OrderedDict()

    def conv(name, o, i, k, bias=True):
        sd[f"{name}.weight"] = (rng.standard_normal((o, i, k, k)) * np.sqrt(2.0 / (i * k * k))).astype(np.float32)
        if bias:
            sd[f"{name}.bias"] = (0.1 * rng.standard_normal(o)).astype(np.float32)

    def bn(name, c):
        sd[f"{name}.weight"] = (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[f"{name}.bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[f"{name}.running_mean"] = np.zeros(c, np.float32)
        sd[f"{name}.running_var"] = np.ones(c, np.float32)
        sd[f"{name}.num_batches_tracked"] = np.zeros((), np.int64)

    for name, l, j in block_names():
        if name.startswith("conv"):
            cin = (in_ch if l == 0 else f[l - 1]) if j == 0 else f[l] * (j + 1) + f[l + 1]
            conv(f"{name}.conv1", f[l], cin, 3)
            bn(f"{name}.bn1", f[l])
            conv(f"{name}.conv2", f[l], f[l], 3)
            bn(f"{name}.bn2", f[l])
        else:
            c = f[l]
            sd[f"{name}.up.weight"] = (rng.standard_normal((c, c, 2, 2)) * np.sqrt(1.0 / c)).astype(np.float32)
            sd[f"{name}.up.bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    conv("ca.fc1", (4 * f[0]) // 16, 4 * f[0], 1, bias=False)
    conv("ca.fc2", 4 * f[0], (4 * f[0]) // 16, 1, bias=False)
    conv("ca1.fc1", f[0] // 4, f[0], 1, bias=False)
    conv("ca1.fc2", f[0], f[0] // 4, 1, bias=False)
    conv("conv_final", out_ch, 4 * f[0], 1)
    return sd


def make_batch(seed: int, N: int, H: int, W: int, in_ch: int = 2):
    """SAR-like synthetic batch: clamp(Exp(mean_c),0,0.15) normalised; labels in {0,1,2,3}."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000003))
    mean = np.array([0.0953, 0.0264, 0.05][:in_ch], np.float32)
    std = np.array([0.0427, 0.0215, 0.03][:in_ch], np.float32)
    xs = []
    for _ in range(2):
        raw = rng.exponential(1.0, size=(N, in_ch, H, W)).astype(np.float32) * mean[None, :, None, None]
        xs.append(((np.clip(raw, 0, 0.15) - mean[None, :, None, None]) / std[None, :, None, None]).astype(np.float32))
    mask = rng.choice(4, size=(N, H, W), p=[0.897, 0.024, 0.041, 0.038]).astype(np.int64)
    return xs[0], xs[1], mask
