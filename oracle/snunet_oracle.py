"""torch-CPU functional restatement of the reference SNUNet-ECAM training step.

Follows models/snunet.py:
  :20-29   conv_block_nested.forward   y1=conv1(x); h=relu(bn1(y1)); out=relu(bn2(conv2(h)) + y1)
  :41      up = ConvTranspose2d(C, C, 2, stride=2)
  :58-62   ChannelAttention            sigmoid(fc2(relu(fc1(avg))) + fc2(relu(fc1(max))))
  :118-153 SNUNet_ECAM.forward         Siamese encoder, nested decoder, ECAM, conv_final
and training/change_detection_trainer.py:136-177 for the step (forward, CE+Dice, backward, Adam).
Operates on a plain dict of tensors keyed like the reference state_dict; gradients come from
autograd over this restatement (the reference's backward IS autograd).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

DEC_ORDER = [(0, 1), (1, 1), (0, 2), (2, 1), (1, 2), (0, 3), (3, 1), (2, 2), (1, 3), (0, 4)]


def to_torch_state(sd_np: Dict[str, np.ndarray], dtype=torch.float32) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd_np.items():
        t = torch.from_numpy(np.array(v))
        out[k] = t.to(dtype) if t.is_floating_point() else t.clone()
    return out


def _block(sd, name, x, training, q, rm=None):
    """rm (tests only): (mask_h, mask_out) boolean tensors that REPLACE the two ReLU sign tests (see snunet_forward)."""
    def act(v, i):
        return v * rm[i].to(v.dtype) if rm is not None else F.relu(v)
    def bn(t, tag):
        return F.batch_norm(t, sd[f"{name}.{tag}.running_mean"], sd[f"{name}.{tag}.running_var"],
                            sd[f"{name}.{tag}.weight"], sd[f"{name}.{tag}.bias"], training, 0.1, 1e-5)
    y1 = q(F.conv2d(x, q(sd[f"{name}.conv1.weight"]), sd[f"{name}.conv1.bias"], padding=1))
    if training:
        sd[f"{name}.bn1.num_batches_tracked"] += 1
        sd[f"{name}.bn2.num_batches_tracked"] += 1
    h = q(act(bn(y1, "bn1"), 0))
    y2 = q(F.conv2d(h, q(sd[f"{name}.conv2.weight"]), sd[f"{name}.conv2.bias"], padding=1))
    return q(act(bn(y2, "bn2") + y1, 1))


def _up(sd, name, x, q):
    return q(F.conv_transpose2d(x, q(sd[f"{name}.up.weight"]), sd[f"{name}.up.bias"], stride=2))


def _ca(sd, name, x):
    avg = F.adaptive_avg_pool2d(x, 1)
    mx = F.adaptive_max_pool2d(x, 1)
    def mlp(t):
        return F.conv2d(F.relu(F.conv2d(t, sd[f"{name}.fc1.weight"])), sd[f"{name}.fc2.weight"])
    return torch.sigmoid(mlp(avg) + mlp(mx))


def snunet_forward(sd: Dict[str, torch.Tensor], xA: torch.Tensor, xB: torch.Tensor, training: bool = True,
                   quant: Optional[Callable] = None, relu_masks: Optional[dict] = None, tap: Optional[dict] = None) -> torch.Tensor:
    """`quant` (optional) rounds stored activations / conv weights, to emulate bf16 storage.
    relu_masks (tests only): {("enc", l, 0|1) / ("dec", l, j): (mask_h, mask_out)} replaces the ReLU sign tests by given masks, so
    that gradient comparisons with another implementation are immune to sign flips of pre-activations that are zero to rounding
    error; tap (tests only) collects each block's (h-less) output under the same keys."""
    rmk = (lambda key: relu_masks[key]) if relu_masks is not None else (lambda key: None)
    q = quant or (lambda t: t)
    X = {}
    for br, x in (("A", q(xA)), ("B", q(xB))):           # snunet.py:120-130 (A first: running-stat order)
        for l in range(5):
            if l == 4 and br == "A":
                continue                                 # snunet.py:124 (commented out)
            inp = x if l == 0 else F.max_pool2d(X[(l - 1, br)], 2, 2)
            X[(l, br)] = _block(sd, f"conv{l}_0", inp, training, q, rmk(("enc", l, 0 if br == "A" else 1)))
            if tap is not None:
                tap[("enc", l, 0 if br == "A" else 1)] = X[(l, br)]
    for (l, j) in DEC_ORDER:                             # snunet.py:132-144
        below = X[(l + 1, "B")] if j == 1 else X[(l + 1, j - 1)]
        cat = [X[(l, "A")], X[(l, "B")]] + [X[(l, k)] for k in range(1, j)] + [_up(sd, f"Up{l + 1}_{j - 1}", below, q)]
        X[(l, j)] = _block(sd, f"conv{l}_{j}", torch.cat(cat, 1), training, q, rmk(("dec", l, j)))
        if tap is not None:
            tap[("dec", l, j)] = X[(l, j)]
    outs = [X[(0, j)] for j in range(1, 5)]
    out = torch.cat(outs, 1)                             # :146
    intra = torch.sum(torch.stack(outs), dim=0)          # :148
    ca1 = _ca(sd, "ca1", intra)                          # :149
    out = _ca(sd, "ca", out) * (out + ca1.repeat(1, 4, 1, 1))   # :150
    return F.conv2d(out, sd["conv_final.weight"], sd["conv_final.bias"])  # :151


def ce_dice_torch(logits: torch.Tensor, mask: torch.Tensor, weights, ignore_index: int = 3) -> torch.Tensor:
    """utilities/bce_and_dice.py:18-24 with utilities/dice.py:111-137, as differentiable torch ops."""
    N, C = logits.shape[:2]
    valid = (mask != ignore_index)
    yd = mask * valid
    t = torch.zeros_like(logits).scatter_(1, yd.unsqueeze(1), 1.0) + 1e-6
    p = F.softmax(logits, dim=1)
    I = torch.sum(p * t, (1, 2, 3))
    S = torch.sum(p + t, (1, 2, 3))
    dice = torch.mean(1.0 - 2.0 * I / (S + 1e-6))
    ce = F.cross_entropy(logits, mask, weight=torch.as_tensor(weights, dtype=logits.dtype), ignore_index=ignore_index)
    return dice + ce


PARAM_SUFFIXES = (".weight", ".bias")


def param_names(sd) -> list:
    return [k for k in sd if k.endswith(PARAM_SUFFIXES)]


def train_step(sd: Dict[str, torch.Tensor], xA, xB, mask, weights=(1.0, 1.0, 1.0), quant=None, relu_masks=None):
    """One forward + loss + backward. Returns (loss, logits, grads dict). Updates BN running stats in sd."""
    names = param_names(sd)
    for k in names:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    logits = snunet_forward(sd, xA, xB, True, quant, relu_masks)
    loss = ce_dice_torch(logits.float(), mask, weights)
    loss.backward()
    grads = {k: sd[k].grad.detach().clone() for k in names}
    for k in names:
        sd[k].requires_grad_(False)
    return loss.detach(), logits.detach(), grads


def adam_step(sd, grads, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam semantics (change_detection_trainer.py:52-54)."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    for k, g in grads.items():
        if weight_decay:
            g = g + weight_decay * sd[k]
        m = state.setdefault(("m", k), torch.zeros_like(g))
        v = state.setdefault(("v", k), torch.zeros_like(g))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
        sd[k].data.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)
