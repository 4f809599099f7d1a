"""CPU restatement (torch functional ops on a plain state dict) of the reference's FC-Siam-conc / FC-Siam-diff.
TEST INFRASTRUCTURE ONLY - never imported by the product.

Follows /root/reference/models/siam_conc.py:95-177 and models/siam_diff.py:95-173 line by line:
  encoder stages (16,16)(32,32)(64,64,64)(128,128,128) of dropout2d(relu(bn(conv3x3))) + maxpool, run on x1 then x2
  with the SAME modules (running statistics are updated twice, x1 first);
  decoder: upconvK = ConvTranspose2d(k3,s2,p1,op1) of the previous stage (stage 4: of the SECOND image's pooled
  feature, siam_conc.py:148 / siam_diff.py:141-144), ReplicationPad2d to the skip size (zero-width for even sizes),
  cat(up, skip_1, skip_2) [conc] or cat(up, |skip_1 - skip_2|) [diff], ConvTranspose2d(k3,s1,p1) stacks,
  Softmax(dim=1) [conc, :177] / LogSoftmax(dim=1) [diff, :173].
Dropout2d is made explicit: `masks[name]` is the [N,C] keep/(1-p) factor of one execution (encoder names carry
the branch suffix `_1` / `_2`); `draw_masks_like_torch` reproduces aten's feature-dropout noise stream so that the
golden generator can pin this restatement against the unmodified reference in train mode with p=0.2.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

ENC = [("11", 16), ("12", 16), ("21", 32), ("22", 32), ("31", 64), ("32", 64), ("33", 64), ("41", 128), ("42", 128), ("43", 128)]
STAGE_LAST = {"12": 1, "22": 2, "33": 3, "43": 4}
DEC = [("43d", 128), ("42d", 128), ("41d", 64), ("33d", 64), ("32d", 64), ("31d", 32), ("22d", 32), ("21d", 16), ("12d", 16)]
UP_BEFORE = {"43d": "upconv4", "33d": "upconv3", "22d": "upconv2", "12d": "upconv1"}
SKIP_OF = {"43d": "43", "33d": "33", "22d": "22", "12d": "12"}
BN_EPS, BN_MOM = 1e-5, 0.1


def exec_order():
    """(mask name, channels) of every Dropout2d execution in forward order (siam_conc.py:99-175)."""
    out = [(f"{n}_1", c) for n, c in ENC] + [(f"{n}_2", c) for n, c in ENC] + list(DEC)
    return out


def make_state(seed: int, in_ch: int = 2, n_cls: int = 3, kind: str = "conc") -> "OrderedDict[str, np.ndarray]":
    """Deterministic state dict with the reference's key order/shapes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()

    def conv(name, o, i):          # nn.Conv2d weight (O, I, 3, 3)
        sd[f"{name}.weight"] = (rng.standard_normal((o, i, 3, 3)) * np.sqrt(2.0 / (i * 9))).astype(np.float32)
        sd[f"{name}.bias"] = (0.1 * rng.standard_normal(o)).astype(np.float32)

    def convt(name, i, o):         # nn.ConvTranspose2d weight (I, O, 3, 3)
        sd[f"{name}.weight"] = (rng.standard_normal((i, o, 3, 3)) * np.sqrt(2.0 / (i * 9))).astype(np.float32)
        sd[f"{name}.bias"] = (0.1 * rng.standard_normal(o)).astype(np.float32)

    def bn(name, c):
        sd[f"{name}.weight"] = (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[f"{name}.bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[f"{name}.running_mean"] = np.zeros(c, np.float32)
        sd[f"{name}.running_var"] = np.ones(c, np.float32)
        sd[f"{name}.num_batches_tracked"] = np.zeros((), np.int64)

    cin = in_ch
    for n, c in ENC:
        conv(f"conv{n}", c, cin)
        bn(f"bn{n}", c)
        cin = c
    mult = 3 if kind == "conc" else 2
    prev = 128
    for n, c in DEC:
        if n in UP_BEFORE:
            convt(UP_BEFORE[n], prev, prev)
            convt(f"conv{n}", prev * mult, c)
        else:
            convt(f"conv{n}", prev, c)
        bn(f"bn{n}", c)
        prev = c
    convt("conv11d", 16, n_cls)
    return sd


def to_torch_state(sd_np) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.array(v)).clone() for k, v in sd_np.items()}


def draw_masks_like_torch(seed: int, N: int, p: float = 0.2) -> Dict[str, torch.Tensor]:
    """aten::feature_dropout: noise = empty([N,C,1,1]).bernoulli_(1-p).div_(1-p), one draw per execution, in order."""
    torch.manual_seed(seed)
    masks = {}
    for name, c in exec_order():
        masks[name] = (torch.empty(N, c, 1, 1).bernoulli_(1 - p) / (1 - p)).reshape(N, c)
    return masks


def _bn(sd, name, y, training):
    if training:
        # F.batch_norm updates running stats in place (momentum 0.1, unbiased variance)
        out = F.batch_norm(y, sd[f"{name}.running_mean"], sd[f"{name}.running_var"], sd[f"{name}.weight"], sd[f"{name}.bias"],
                           True, BN_MOM, BN_EPS)
        sd[f"{name}.num_batches_tracked"] += 1
        return out
    return F.batch_norm(y, sd[f"{name}.running_mean"], sd[f"{name}.running_var"], sd[f"{name}.weight"], sd[f"{name}.bias"],
                        False, BN_MOM, BN_EPS)


def _do(x, masks, key):
    if masks is None or key not in masks:
        return x
    return x * masks[key][:, :, None, None]


def siam_forward(sd: Dict[str, torch.Tensor], x1, x2, kind: str = "conc", training: bool = True,
                 masks: Optional[Dict[str, torch.Tensor]] = None, tap: Optional[dict] = None,
                 relu_masks: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """Returns the model output: probabilities (conc) or log-probabilities (diff), [N,K,H,W].
    relu_masks (tests only): per-execution boolean masks that REPLACE the ReLU sign test (v * mask instead of relu(v)); with the
    masks of another implementation's activations this makes gradient comparisons immune to sign flips of pre-activations
    that are zero to rounding error (a ReLU mask flip changes the gradient by a finite amount)."""
    def act(v, key):
        return v * relu_masks[key].to(v.dtype) if relu_masks is not None else F.relu(v)

    skips = {}
    pooled_last = None
    for br, x in ((1, x1), (2, x2)):
        h = x
        for n, _ in ENC:                                   # siam_conc.py:99-144
            h = _do(act(_bn(sd, f"bn{n}", F.conv2d(h, sd[f"conv{n}.weight"], sd[f"conv{n}.bias"], padding=1), training), f"{n}_{br}"),
                    masks, f"{n}_{br}")
            if tap is not None:
                tap[f"{n}_{br}"] = h
            if n in STAGE_LAST:
                skips[(n, br)] = h
                h = F.max_pool2d(h, 2, 2)
        pooled_last = h                                    # x4p_2 after the second pass (:148)
    h = pooled_last
    for n, _ in DEC:                                       # siam_conc.py:146-175
        if n in UP_BEFORE:
            u = UP_BEFORE[n]
            h = F.conv_transpose2d(h, sd[f"{u}.weight"], sd[f"{u}.bias"], stride=2, padding=1, output_padding=1)
            s1, s2 = skips[(SKIP_OF[n], 1)], skips[(SKIP_OF[n], 2)]
            h = F.pad(h, (0, s1.shape[3] - h.shape[3], 0, s1.shape[2] - h.shape[2]), mode="replicate")
            h = torch.cat((h, s1, s2), 1) if kind == "conc" else torch.cat((h, torch.abs(s1 - s2)), 1)
        h = _do(act(_bn(sd, f"bn{n}", F.conv_transpose2d(h, sd[f"conv{n}.weight"], sd[f"conv{n}.bias"], padding=1), training), n),
                masks, n)
        if tap is not None:
            tap[n] = h
    z = F.conv_transpose2d(h, sd["conv11d.weight"], sd["conv11d.bias"], padding=1)
    return F.softmax(z, 1) if kind == "conc" else F.log_softmax(z, 1)


def train_step(sd, x1, x2, mask, kind="conc", class_weights=(1.0, 1.0, 1.0), masks=None, relu_masks=None):
    """forward (train-mode BN) + CE+Dice on the model OUTPUT (the reference applies the criterion to the softmax /
    log-softmax output: change_detection_trainer.py:136-170, utilities.py:342-347) + autograd backward."""
    from .snunet_oracle import ce_dice_torch
    names = [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work = dict(sd)
    work.update(leaves)
    out = siam_forward(work, x1, x2, kind, True, masks, None, relu_masks)
    # running stats / num_batches_tracked were updated in place: `work` shares those tensor objects with `sd`
    loss = ce_dice_torch(out, mask, class_weights)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return loss.detach(), out.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}
