"""numpy restatement of the reference CE+Dice loss (forward, closed-form gradient, argmax).

Follows, line by line:
  utilities/bce_and_dice.py:18-24   loss = dice(preds, lbl) + CrossEntropyLoss(weight, ignore_index)(preds, lbl)
  utilities/dice.py:111-119         target_cl = target * (target != ignore_index)   (ignored pixels -> class 0)
  utilities/dice.py:57-59           one_hot = zeros(int64).scatter_(1, y, 1.0) + 1e-6 -> float32 {1e-6, 1+1e-6}
  utilities/dice.py:127-137         softmax over C; per-sample sums over (C,H,W); mean_n(1 - 2I/(S + 1e-6))
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np


def _softmax(z: np.ndarray) -> np.ndarray:
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m)
    return e / e.sum(axis=1, keepdims=True)


def ce_dice(logits: np.ndarray, labels: np.ndarray, weights, ignore_index: int = 3, dtype=np.float64):
    """Returns dict(loss, dice, ce, dlogits, argmax). logits [N,C,H,W], labels [N,H,W] int64."""
    z = logits.astype(dtype)
    N, C, H, W = z.shape
    y = labels.astype(np.int64)
    w = np.asarray(weights, dtype=dtype)
    p = _softmax(z)
    valid = (y != ignore_index)
    yd = y * valid                                             # dice.py:119
    eps32 = np.float32(1e-6)
    t = np.full((N, C, H, W), dtype(eps32), dtype=dtype)       # dice.py:59 (+eps in fp32)
    one = dtype(np.float32(1.0) + eps32)
    n_i, h_i, w_i = np.meshgrid(np.arange(N), np.arange(H), np.arange(W), indexing="ij")
    t[n_i, yd, h_i, w_i] = one
    I = (p * t).sum(axis=(1, 2, 3))                            # dice.py:133
    S = (p + t).sum(axis=(1, 2, 3))                            # dice.py:134
    dice = np.mean(1.0 - 2.0 * I / (S + 1e-6))                 # dice.py:136-137
    # nn.CrossEntropyLoss(weight, ignore_index): sum_valid w[y] * (-log p_y) / sum_valid w[y]
    ysafe = np.where(valid, y, 0)
    logp = np.log(p)
    lp_y = logp[n_i, ysafe, h_i, w_i]
    wy = w[ysafe] * valid
    den = wy.sum()
    with np.errstate(invalid="ignore", divide="ignore"):
        ce = (-(wy * lp_y).sum()) / den                        # NaN when every pixel is ignored (as torch)
    # gradient: dL/dp = a_n t + b_n ; dL/dz = p * (g - sum_c g p)  (+ CE term)
    a = (-2.0 / (N * (S + 1e-6)))[:, None, None, None]
    b = (2.0 * I / (N * (S + 1e-6) ** 2))[:, None, None, None]
    g = a * t + b
    dz = p * (g - (g * p).sum(axis=1, keepdims=True))
    onehot_y = np.zeros_like(p)
    onehot_y[n_i, ysafe, h_i, w_i] = 1.0
    with np.errstate(invalid="ignore", divide="ignore"):
        dz = dz + (p - onehot_y) * (wy / den)[:, None, :, :]
    return {"loss": dice + ce, "dice": dice, "ce": ce, "dlogits": dz, "argmax": np.argmax(logits, axis=1).astype(np.uint8)}
