"""CPU restatement (torch functional ops on a plain state dict) of the reference's FloodViT path.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/models/vision_transformer.py line by line - ViT.forward :139-156 (patchify 'b c (h p1) (w p2) ->
b (h w) (p1 p2 c)', LayerNorm, Linear, LayerNorm, cls token, + pos_embedding, transformer, drop the cls token),
Transformer.forward :84-89 (x = attn(x) + x; x = ff(x) + x; final LayerNorm), Attention.forward :53-66 (pre-norm, bias-free
qkv, softmax(q k^T d^-1/2) v, output projection), FeedForward :19-32 (LayerNorm, Linear, exact GELU, Linear) - and
models/model_utilities.py:80-94 (FinetunerSegmentation.forward: tokens -> [B, dim, G, G] -> bilinear upsample to 224 ->
1x1 conv head).  State-dict keys are those of FinetunerSegmentation (`model.*` encoder, `head.*`).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


def make_state(seed: int, dim: int, depth: int, heads: int, mlp_dim: int, channels: int = 6, n_cls: int = 3, tokens: int = 197,
               dim_head: int = 64) -> "OrderedDict[str, np.ndarray]":
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    pd, inner = 256 * channels, heads * dim_head

    def ln(name, c):
        sd[f"{name}.weight"] = (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[f"{name}.bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)

    def lin(name, o, i, bias=True):
        sd[f"{name}.weight"] = (rng.standard_normal((o, i)) * np.sqrt(1.0 / i)).astype(np.float32)
        if bias:
            sd[f"{name}.bias"] = (0.1 * rng.standard_normal(o)).astype(np.float32)

    sd["model.pos_embedding"] = (0.5 * rng.standard_normal((1, tokens, dim))).astype(np.float32)
    sd["model.cls_token"] = (0.5 * rng.standard_normal((1, 1, dim))).astype(np.float32)
    ln("model.to_patch_embedding.1", pd)
    lin("model.to_patch_embedding.2", dim, pd)
    ln("model.to_patch_embedding.3", dim)
    ln("model.transformer.norm", dim)
    for l in range(depth):
        p = f"model.transformer.layers.{l}"
        ln(f"{p}.0.norm", dim)
        lin(f"{p}.0.to_qkv", 3 * inner, dim, bias=False)
        lin(f"{p}.0.to_out.0", dim, inner)
        ln(f"{p}.1.net.0", dim)
        lin(f"{p}.1.net.1", mlp_dim, dim)
        lin(f"{p}.1.net.4", dim, mlp_dim)
    sd["head.weight"] = (rng.standard_normal((n_cls, dim, 1, 1)) * np.sqrt(1.0 / dim)).astype(np.float32)
    sd["head.bias"] = (0.1 * rng.standard_normal(n_cls)).astype(np.float32)
    return sd


def make_batch(seed: int, N: int, H: int = 224, W: int = 224, channels: int = 6):
    """Segmentation batch: image = cat(post, pre1, pre2) (segmentation_trainer.py:138-144), SAR-like values, labels in {0..3}."""
    rng = np.random.Generator(np.random.PCG64(seed + 2000003))
    mean = np.array([0.0953, 0.0264] * 3, np.float32)[:channels]
    std = np.array([0.0427, 0.0215] * 3, np.float32)[:channels]
    raw = rng.exponential(1.0, size=(N, channels, H, W)).astype(np.float32) * mean[None, :, None, None]
    img = ((np.clip(raw, 0, 0.15) - mean[None, :, None, None]) / std[None, :, None, None]).astype(np.float32)
    mask = rng.choice(4, size=(N, H, W), p=[0.897, 0.024, 0.041, 0.038]).astype(np.int64)
    return img, mask


def to_torch_state(sd_np) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.array(v)).clone() for k, v in sd_np.items()}


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{name}.weight"], sd[f"{name}.bias"], 1e-5)


def vit_tokens(sd, img, heads: int, dim_head: int = 64, pre: str = "model.", out_indices=None):
    """out_indices (optional): also return the residual stream (cls token dropped) after those block counts; an index equal to
    the depth yields the final-norm output."""
    B, C, H, W = img.shape
    gh, gw = H // 16, W // 16
    x = img.view(B, C, gh, 16, gw, 16).permute(0, 2, 4, 3, 5, 1).reshape(B, gh * gw, 256 * C)       # :122
    pe = f"{pre}to_patch_embedding"
    x = _ln(sd, f"{pe}.3", F.linear(_ln(sd, f"{pe}.1", x), sd[f"{pe}.2.weight"], sd[f"{pe}.2.bias"]))   # :123-125
    n = x.shape[1]
    x = torch.cat((sd[f"{pre}cls_token"].expand(B, -1, -1), x), dim=1)                          # :142-143
    x = x + sd[f"{pre}pos_embedding"][:, : n + 1]                                              # :144
    depth = 1 + max(int(k.split(".")[3 if pre else 2]) for k in sd if k.startswith(f"{pre}transformer.layers."))
    inner = heads * dim_head
    taps = {}
    for l in range(depth):
        pa, pf = f"{pre}transformer.layers.{l}.0", f"{pre}transformer.layers.{l}.1.net"
        y = _ln(sd, f"{pa}.norm", x)                                                           # :54
        q, k, v = [t.view(B, n + 1, heads, dim_head).transpose(1, 2) for t in F.linear(y, sd[f"{pa}.to_qkv.weight"]).chunk(3, dim=-1)]
        attn = torch.softmax(q @ k.transpose(-1, -2) * dim_head ** -0.5, dim=-1)                # :59-61
        o = (attn @ v).transpose(1, 2).reshape(B, n + 1, inner)                                # :64-65
        x = F.linear(o, sd[f"{pa}.to_out.0.weight"], sd[f"{pa}.to_out.0.bias"]) + x             # :86
        y = _ln(sd, f"{pf}.0", x)
        y = F.linear(F.gelu(F.linear(y, sd[f"{pf}.1.weight"], sd[f"{pf}.1.bias"])), sd[f"{pf}.4.weight"], sd[f"{pf}.4.bias"])
        x = y + x                                                                              # :87
        taps[l + 1] = x[:, 1:]
    x = _ln(sd, f"{pre}transformer.norm", x)                                                   # :89
    if out_indices is not None:
        taps[depth] = x[:, 1:]
        return [taps[i] for i in out_indices]
    return x[:, 1:]                                                                            # :152


def floodvit_forward(sd, img, heads: int, dim_head: int = 64, out_size: int = 224) -> torch.Tensor:
    tok = vit_tokens(sd, img, heads, dim_head)
    B, n, D = tok.shape
    G = int(round(n ** 0.5))
    x = tok.view(B, G, G, D).permute(0, 3, 1, 2)                                               # model_utilities.py:87
    if "head.deconv1.weight" in sd:                                                            # configs["decoder"]: no interpolation (:88), Decoder.forward (:36-48)
        x = F.relu(F.conv_transpose2d(x, sd["head.deconv1.weight"], sd["head.deconv1.bias"], stride=2, padding=1))
        x = F.interpolate(x, scale_factor=2)                                                   # nn.Upsample(scale_factor=2): nearest
        x = F.relu(F.conv_transpose2d(x, sd["head.deconv2.weight"], sd["head.deconv2.bias"], stride=2, padding=1))
        return F.conv_transpose2d(x, sd["head.deconv3.weight"], sd["head.deconv3.bias"], stride=2, padding=1)
    x = F.interpolate(x, size=(out_size, out_size), mode="bilinear", align_corners=False)      # :89-91
    if "head.0.weight" in sd:                                                                  # configs["mlp"]: Conv1x1 -> ReLU -> Conv1x1 (:60-65)
        x = F.relu(F.conv2d(x, sd["head.0.weight"], sd["head.0.bias"]))
        return F.conv2d(x, sd["head.2.weight"], sd["head.2.bias"])
    return F.conv2d(x, sd["head.weight"], sd["head.bias"])                                     # :93


def make_state_mlp(seed: int, dim: int, depth: int, heads: int, mlp_dim: int, hidden: int = 512, n_cls: int = 3):
    """State dict of FinetunerSegmentation(configs mlp=True): head = Sequential(Conv2d(dim, 512, 1), ReLU, Conv2d(512, n_cls, 1))."""
    sd = make_state(seed, dim, depth, heads, mlp_dim)
    del sd["head.weight"], sd["head.bias"]
    rng = np.random.Generator(np.random.PCG64(seed + 17))
    sd["head.0.weight"] = (rng.standard_normal((hidden, dim, 1, 1)) * np.sqrt(2.0 / dim)).astype(np.float32)
    sd["head.0.bias"] = (0.1 * rng.standard_normal(hidden)).astype(np.float32)
    sd["head.2.weight"] = (rng.standard_normal((n_cls, hidden, 1, 1)) * np.sqrt(1.0 / hidden)).astype(np.float32)
    sd["head.2.bias"] = (0.1 * rng.standard_normal(n_cls)).astype(np.float32)
    return sd


def make_state_decoder(seed: int, depth: int, heads: int, mlp_dim: int, n_cls: int = 3):
    """State dict of FinetunerSegmentation(configs decoder=True): head = Decoder (deconv 1024->128->64->n_cls, all k4 s2 p1); the
    reference hard-wires the first deconvolution to 1024 input channels, so the encoder width is 1024."""
    dim = 1024
    sd = make_state(seed, dim, depth, heads, mlp_dim)
    del sd["head.weight"], sd["head.bias"]
    rng = np.random.Generator(np.random.PCG64(seed + 23))
    for name, ci, co in (("deconv1", 1024, 128), ("deconv2", 128, 64), ("deconv3", 64, n_cls)):
        sd[f"head.{name}.weight"] = (rng.standard_normal((ci, co, 4, 4)) * np.sqrt(2.0 / (ci * 4))).astype(np.float32)
        sd[f"head.{name}.bias"] = (0.1 * rng.standard_normal(co)).astype(np.float32)
    return sd


def train_step(sd, img, mask, heads: int, class_weights=(1.0, 1.0, 1.0)):
    from .snunet_oracle import ce_dice_torch
    names = list(sd.keys())
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    logits = floodvit_forward(leaves, img, heads, out_size=img.shape[-1])
    loss = ce_dice_torch(logits, mask, class_weights)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return loss.detach(), logits.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}
