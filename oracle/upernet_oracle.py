"""CPU restatement of HF `UperNetHead` (transformers/models/upernet/modeling_upernet.py:  UperNetConvModule = conv(bias=False) ->
BatchNorm2d -> ReLU; UperNetPyramidPoolingModule = AdaptiveAvgPool2d(s) -> 1x1 ConvModule -> bilinear resize, s in (1,2,3,6);
UperNetHead.forward = lateral 1x1 ConvModules, PSP bottleneck on the last feature, top-down adds, 3x3 FPN ConvModules, concat,
3x3 fpn_bottleneck, 1x1 classifier) and of the FloodViT + UPerNet composition this repo defines.  TEST INFRASTRUCTURE ONLY.

The reference reaches this head only through `models/upernet.py:80` (`UperNetForSemanticSegmentation`, transformers==4.31.0 pinned
at requirements.txt:27, ConvNeXt / Swin backbones with pretrained weights that need network).  BASELINE.json's config "FloodViT
(MAE-ViT-B encoder + UPerNet head)" has no counterpart in the reference; it is DEFINED here as: the
reference ViT encoder (oracle/vit_oracle.py), token maps (cls dropped, 14x14) after blocks `out_indices` (the last one after the
final LayerNorm), HF UperNetHead(hidden 512, pool scales 1/2/3/6) on them, logits resized to the input size (bilinear,
align_corners=False, as UperNetForSemanticSegmentation.forward does).  The head is pinned against the installed HF class
(oracle/make_golden.py); the composition is parity-pinned at the op level only.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import vit_oracle

SCALES = (1, 2, 3, 6)


def make_head_state(seed: int, in_channels, hidden: int = 512, n_cls: int = 3, prefix: str = "decode_head.") -> "OrderedDict[str, np.ndarray]":
    """HF UperNetHead state dict (key order of the installed class)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()

    def cbr(name, o, i, k):
        sd[f"{prefix}{name}.conv.weight"] = (rng.standard_normal((o, i, k, k)) * np.sqrt(2.0 / (i * k * k))).astype(np.float32)
        sd[f"{prefix}{name}.batch_norm.weight"] = (1.0 + 0.1 * rng.standard_normal(o)).astype(np.float32)
        sd[f"{prefix}{name}.batch_norm.bias"] = (0.1 * rng.standard_normal(o)).astype(np.float32)
        sd[f"{prefix}{name}.batch_norm.running_mean"] = np.zeros(o, np.float32)
        sd[f"{prefix}{name}.batch_norm.running_var"] = np.ones(o, np.float32)
        sd[f"{prefix}{name}.batch_norm.num_batches_tracked"] = np.zeros((), np.int64)

    sd[f"{prefix}classifier.weight"] = (rng.standard_normal((n_cls, hidden, 1, 1)) * np.sqrt(1.0 / hidden)).astype(np.float32)
    sd[f"{prefix}classifier.bias"] = (0.1 * rng.standard_normal(n_cls)).astype(np.float32)
    for i in range(len(SCALES)):
        cbr(f"psp_modules.{i}.1", hidden, in_channels[-1], 1)
    cbr("bottleneck", hidden, in_channels[-1] + len(SCALES) * hidden, 3)
    for i, c in enumerate(in_channels[:-1]):
        cbr(f"lateral_convs.{i}", hidden, c, 1)
    for i in range(len(in_channels) - 1):
        cbr(f"fpn_convs.{i}", hidden, hidden, 3)
    cbr("fpn_bottleneck", hidden, len(in_channels) * hidden, 3)
    return sd


def _conv_module(sd, name, x, training, pad, relu_masks=None, tap=None):
    y = F.conv2d(x, sd[f"{name}.conv.weight"], None, padding=pad)
    y = F.batch_norm(y, sd[f"{name}.batch_norm.running_mean"], sd[f"{name}.batch_norm.running_var"], sd[f"{name}.batch_norm.weight"],
                     sd[f"{name}.batch_norm.bias"], training, 0.1, 1e-5)
    if training:
        sd[f"{name}.batch_norm.num_batches_tracked"] += 1
    if tap is not None:
        tap[name] = y.detach()
    return y * relu_masks[name].to(y.dtype) if relu_masks is not None else F.relu(y)


def upernet_head(sd: Dict[str, torch.Tensor], feats: List[torch.Tensor], training: bool = True, prefix: str = "decode_head.",
                 relu_masks=None, tap=None) -> torch.Tensor:
    """feats: NCHW maps, coarsest last.  Returns the logits at the resolution of feats[0].
    relu_masks (tests only): per-module boolean masks that REPLACE the ReLU sign test (see oracle/siam_oracle.py:siam_forward);
    tap: receives every module's pre-ReLU map."""
    p = prefix

    def _cbr(sd, name, x, training, pad):
        return _conv_module(sd, name, x, training, pad, relu_masks, tap)
    laterals = [_cbr(sd, f"{p}lateral_convs.{i}", feats[i], training, 0) for i in range(len(feats) - 1)]
    x = feats[-1]
    psp = [x]
    for i, s in enumerate(SCALES):
        o = _cbr(sd, f"{p}psp_modules.{i}.1", F.adaptive_avg_pool2d(x, s), training, 0)
        psp.append(F.interpolate(o, size=x.shape[2:], mode="bilinear", align_corners=False))
    laterals.append(_cbr(sd, f"{p}bottleneck", torch.cat(psp, 1), training, 1))
    n = len(laterals)
    for i in range(n - 1, 0, -1):
        laterals[i - 1] = laterals[i - 1] + F.interpolate(laterals[i], size=laterals[i - 1].shape[2:], mode="bilinear", align_corners=False)
    outs = [_cbr(sd, f"{p}fpn_convs.{i}", laterals[i], training, 1) for i in range(n - 1)] + [laterals[-1]]
    for i in range(n - 1, 0, -1):
        outs[i] = F.interpolate(outs[i], size=outs[0].shape[2:], mode="bilinear", align_corners=False)
    out = _cbr(sd, f"{p}fpn_bottleneck", torch.cat(outs, 1), training, 1)
    return F.conv2d(out, sd[f"{p}classifier.weight"], sd[f"{p}classifier.bias"])


def make_state(seed: int, dim: int, depth: int, heads: int, mlp_dim: int, hidden: int = 512, n_cls: int = 3):
    sd = vit_oracle.make_state(seed, dim, depth, heads, mlp_dim)
    del sd["head.weight"], sd["head.bias"]
    sd.update(make_head_state(seed + 1, [dim] * 4, hidden, n_cls))
    return sd


def forward(sd, img, heads: int, out_indices, training: bool = True, relu_masks=None, tap=None) -> torch.Tensor:
    taps = vit_oracle.vit_tokens(sd, img, heads, out_indices=list(out_indices))
    B, n, D = taps[0].shape
    G = int(round(n ** 0.5))
    feats = [t.reshape(B, G, G, D).permute(0, 3, 1, 2) for t in taps]
    logits = upernet_head(sd, feats, training, relu_masks=relu_masks, tap=tap)
    return F.interpolate(logits, size=img.shape[2:], mode="bilinear", align_corners=False)


def train_step(sd, img, mask, heads: int, out_indices, class_weights=(1.0, 1.0, 1.0), relu_masks=None, tap=None):
    from .snunet_oracle import ce_dice_torch
    names = [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work = dict(sd)
    work.update(leaves)
    logits = forward(work, img, heads, out_indices, True, relu_masks, tap)
    loss = ce_dice_torch(logits, mask, class_weights)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return loss.detach(), logits.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}
