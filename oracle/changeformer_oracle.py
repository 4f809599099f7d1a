"""CPU restatement (torch functional ops on a plain state dict) of the reference's ChangeFormerV6.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/models/changeformer.py line by line:
  OverlapPatchEmbed.forward :285-292   conv 7x7 (stride 4 in stage 1, 2 afterwards; padding 3) -> tokens -> LayerNorm(eps 1e-5)
  Attention.forward         :186-208   q = Linear(x); x_ = LayerNorm(conv_{k=s=sr}(x)) if sr > 1; k, v = Linear(x_);
                                       softmax(q k^T d^-1/2) v; Linear                (attn / proj dropout: identity here)
  Mlp / DWConv              :90-96,126-133   fc2(gelu(dwconv3x3(fc1(x))))            (dropout: identity here)
  Block.forward             :244-248   x += attn(LN_1e-6(x)); x += mlp(LN_1e-6(x))    (DropPath: identity here)
  EncoderTransformer_v3     :430-465   4 stages of (patch embed, blocks [3,3,4,3], LayerNorm(eps 1e-6)) -> NCHW feature maps
  DecoderTransformer_v3     :568-641   per scale: Linear embed of both dates, conv_diff(cat) (+ bilinear x2 of the coarser scale),
                                       side predictions, bilinear resize to the 1/4 scale, 1x1 fuse + BN, ConvT4x4s2 + residual
                                       block (x2), 3x3 classifier, Sigmoid on all five outputs when decoder_softmax
  ChangeFormerV6.forward    :666-676   shared encoder on both dates, decoder
The stochastic layers (Dropout 0.1, attention dropout 0.1, DropPath 0.1 - :652-654) cannot share an RNG stream with another
implementation, so parity runs use p = 0 (train-mode BatchNorm statistics are kept); SURVEY.md section 7.3 item 7.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

EMBED_DIMS, DEPTHS, HEADS, SR = [64, 128, 320, 512], [3, 3, 4, 3], [1, 2, 4, 8], [8, 4, 2, 1]


# ---- stateless RNG of the product's stochastic layers, restated (bit for bit: csrc/cformer.cu cf_mix64 / cf_key / cf_keep) ------------
def _mix64(x):
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def keep_factors(n: int, p: float, seed: int, step: int, site: int) -> torch.Tensor:
    """fp32 [n]: 0 with probability p, 1/(1-p) otherwise - the factor the product applies to element i at (seed, step, site)."""
    with np.errstate(over="ignore"):
        k0 = (int(seed) ^ ((int(step) & 0xFFFFFFFF) << 32) ^ ((int(site) & 0xFFFFFFFF) * 0x632BE59BD9B4E019)) & ((1 << 64) - 1)
        key = _mix64(np.array([k0], dtype=np.uint64))[0]
        r = _mix64(key + np.arange(n, dtype=np.uint64))
    u = (r >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return torch.from_numpy(np.where(u >= np.float32(p), np.float32(1.0) / (np.float32(1.0) - np.float32(p)), np.float32(0.0)).astype(np.float32))


class Stoch:
    """Explicit stochastic-layer masks for parity runs against the product: the product's own RNG (keep_factors) for the element
    dropouts, its DropPath factors `dp` [blocks][2][2B] as drawn on the device.  Needs the two dates batched as cat(x1, x2)."""

    def __init__(self, seed, step, p_drop, p_attn, dp):
        self.seed, self.step, self.p_drop, self.p_attn, self.dp = seed, step, p_drop, p_attn, dp

    def elem(self, t, p, site):
        return t if p <= 0 else t * keep_factors(t.numel(), p, self.seed, self.step, site).view(t.shape)


def make_state(seed: int, in_ch: int = 2, n_cls: int = 3, embed_dim: int = 256, embed_dims=None, depths=None) -> "OrderedDict[str, np.ndarray]":
    """Deterministic state dict in the reference's key order / shapes (373 entries for the default config)."""
    dims, dep = embed_dims or EMBED_DIMS, depths or DEPTHS
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()

    def w(name, shape, fan_in, bias=None):
        sd[f"{name}.weight"] = (rng.standard_normal(shape) * np.sqrt(1.0 / fan_in)).astype(np.float32)
        if bias is not None:
            sd[f"{name}.bias"] = (0.1 * rng.standard_normal(bias)).astype(np.float32)

    def ln(name, c):
        sd[f"{name}.weight"] = (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[f"{name}.bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)

    def bn(name, c):
        ln(name, c)
        sd[f"{name}.running_mean"] = np.zeros(c, np.float32)
        sd[f"{name}.running_var"] = np.ones(c, np.float32)
        sd[f"{name}.num_batches_tracked"] = np.zeros((), np.int64)

    cin = in_ch
    for s in range(4):
        w(f"Tenc_x2.patch_embed{s + 1}.proj", (dims[s], cin, 7, 7), cin * 49, dims[s])
        ln(f"Tenc_x2.patch_embed{s + 1}.norm", dims[s])
        cin = dims[s]
    for s in range(4):
        C = dims[s]
        for i in range(dep[s]):
            p = f"Tenc_x2.block{s + 1}.{i}"
            ln(f"{p}.norm1", C)
            w(f"{p}.attn.q", (C, C), C, C)
            w(f"{p}.attn.kv", (2 * C, C), C, 2 * C)
            w(f"{p}.attn.proj", (C, C), C, C)
            if SR[s] > 1:
                w(f"{p}.attn.sr", (C, C, SR[s], SR[s]), C * SR[s] * SR[s], C)
                ln(f"{p}.attn.norm", C)
            ln(f"{p}.norm2", C)
            w(f"{p}.mlp.fc1", (4 * C, C), C, 4 * C)
            w(f"{p}.mlp.dwconv.dwconv", (4 * C, 1, 3, 3), 9, 4 * C)
            w(f"{p}.mlp.fc2", (C, 4 * C), 4 * C, C)
        ln(f"Tenc_x2.norm{s + 1}", C)
    E = embed_dim
    for s in (4, 3, 2, 1):
        w(f"TDec_x2.linear_c{s}.proj", (E, dims[s - 1]), dims[s - 1], E)
    for s in (4, 3, 2, 1):
        w(f"TDec_x2.diff_c{s}.0", (E, 2 * E, 3, 3), 2 * E * 9, E)
        bn(f"TDec_x2.diff_c{s}.2", E)
        w(f"TDec_x2.diff_c{s}.3", (E, E, 3, 3), E * 9, E)
    for s in (4, 3, 2, 1):
        w(f"TDec_x2.make_pred_c{s}.0", (n_cls, E, 3, 3), E * 9, n_cls)
        bn(f"TDec_x2.make_pred_c{s}.2", n_cls)
        w(f"TDec_x2.make_pred_c{s}.3", (n_cls, n_cls, 3, 3), n_cls * 9, n_cls)
    w("TDec_x2.linear_fuse.0", (E, 4 * E, 1, 1), 4 * E, E)
    bn("TDec_x2.linear_fuse.1", E)
    w("TDec_x2.convd2x.conv2d", (E, E, 4, 4), E * 4, E)           # ConvTranspose2d: (Cin, Cout, 4, 4)
    w("TDec_x2.dense_2x.0.conv1.conv2d", (E, E, 3, 3), E * 9, E)
    w("TDec_x2.dense_2x.0.conv2.conv2d", (E, E, 3, 3), E * 9, E)
    w("TDec_x2.convd1x.conv2d", (E, E, 4, 4), E * 4, E)
    w("TDec_x2.dense_1x.0.conv1.conv2d", (E, E, 3, 3), E * 9, E)
    w("TDec_x2.dense_1x.0.conv2.conv2d", (E, E, 3, 3), E * 9, E)
    w("TDec_x2.change_probability.conv2d", (n_cls, E, 3, 3), E * 9, n_cls)
    return sd


def to_torch_state(sd_np) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.array(v)).clone() for k, v in sd_np.items()}


def _ln(sd, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{name}.weight"], sd[f"{name}.bias"], eps)


def _lin(sd, name, x):
    return F.linear(x, sd[f"{name}.weight"], sd[f"{name}.bias"])


def _bn(sd, name, x, training):
    out = F.batch_norm(x, sd[f"{name}.running_mean"], sd[f"{name}.running_var"], sd[f"{name}.weight"], sd[f"{name}.bias"], training, 0.1, 1e-5)
    if training:
        sd[f"{name}.num_batches_tracked"] += 1
    return out


def _attention(sd, p, x, H, W, heads, sr, st=None, g=0):          # :186-208
    B, N, C = x.shape
    d = C // heads
    q = _lin(sd, f"{p}.q", x).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    if sr > 1:
        x_ = x.permute(0, 2, 1).reshape(B, C, H, W)
        x_ = F.conv2d(x_, sd[f"{p}.sr.weight"], sd[f"{p}.sr.bias"], stride=sr).reshape(B, C, -1).permute(0, 2, 1)
        x_ = _ln(sd, f"{p}.norm", x_, 1e-5)                       # nn.LayerNorm(dim): default eps (:167)
    else:
        x_ = x
    kv = _lin(sd, f"{p}.kv", x_).reshape(B, -1, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = ((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(dim=-1)
    if st is not None:
        attn = st.elem(attn, st.p_attn, g * 8)                    # attn_drop (:203)
    out = _lin(sd, f"{p}.proj", (attn @ v).transpose(1, 2).reshape(B, N, C))
    return out if st is None else st.elem(out, st.p_drop, g * 8 + 1)   # proj_drop (:206)


def _mlp(sd, p, x, H, W, st=None, g=0):                         # :126-133, :90-96
    B, N, _ = x.shape
    h = _lin(sd, f"{p}.fc1", x)
    Ch = h.shape[-1]
    h = F.conv2d(h.transpose(1, 2).reshape(B, Ch, H, W), sd[f"{p}.dwconv.dwconv.weight"], sd[f"{p}.dwconv.dwconv.bias"], padding=1, groups=Ch)
    h = F.gelu(h.flatten(2).transpose(1, 2))
    if st is not None:
        h = st.elem(h, st.p_drop, g * 8 + 2)                      # drop after the activation (:130)
    out = _lin(sd, f"{p}.fc2", h)
    return out if st is None else st.elem(out, st.p_drop, g * 8 + 3)   # drop after fc2 (:132)


def encoder(sd, x, dims=None, depths=None, st=None) -> List[torch.Tensor]:   # :430-465
    dims, dep = dims or EMBED_DIMS, depths or DEPTHS
    outs = []
    B = x.shape[0]
    g = 0
    for s in range(4):
        pe = f"Tenc_x2.patch_embed{s + 1}"
        x = F.conv2d(x, sd[f"{pe}.proj.weight"], sd[f"{pe}.proj.bias"], stride=4 if s == 0 else 2, padding=3)
        H, W = x.shape[2:]
        t = _ln(sd, f"{pe}.norm", x.flatten(2).transpose(1, 2), 1e-5)   # OverlapPatchEmbed.norm: default eps (:266)
        for i in range(dep[s]):
            p = f"Tenc_x2.block{s + 1}.{i}"
            dp1 = 1.0 if st is None else st.dp[g, 0].view(B, 1, 1)          # drop_path (:246-247)
            dp2 = 1.0 if st is None else st.dp[g, 1].view(B, 1, 1)
            t = t + dp1 * _attention(sd, f"{p}.attn", _ln(sd, f"{p}.norm1", t, 1e-6), H, W, HEADS[s], SR[s], st, g)
            t = t + dp2 * _mlp(sd, f"{p}.mlp", _ln(sd, f"{p}.norm2", t, 1e-6), H, W, st, g)
            g += 1
        t = _ln(sd, f"Tenc_x2.norm{s + 1}", t, 1e-6)
        x = t.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return outs


def decoder(sd, f1, f2, training=True, decoder_softmax=True) -> List[torch.Tensor]:    # :568-641
    D = "TDec_x2"
    n = f1[0].shape[0]
    size1 = f1[0].shape[2:]
    outs, ups, prev = [], [], None
    for s in (4, 3, 2, 1):
        a, b = f1[s - 1], f2[s - 1]
        emb = lambda c: _lin(sd, f"{D}.linear_c{s}.proj", c.flatten(2).transpose(1, 2)).permute(0, 2, 1).reshape(n, -1, c.shape[2], c.shape[3])
        x = torch.cat((emb(a), emb(b)), dim=1)
        x = F.relu(F.conv2d(x, sd[f"{D}.diff_c{s}.0.weight"], sd[f"{D}.diff_c{s}.0.bias"], padding=1))
        x = _bn(sd, f"{D}.diff_c{s}.2", x, training)
        x = F.relu(F.conv2d(x, sd[f"{D}.diff_c{s}.3.weight"], sd[f"{D}.diff_c{s}.3.bias"], padding=1))
        if prev is not None:
            x = x + F.interpolate(prev, scale_factor=2, mode="bilinear")
        p = F.relu(F.conv2d(x, sd[f"{D}.make_pred_c{s}.0.weight"], sd[f"{D}.make_pred_c{s}.0.bias"], padding=1))
        p = _bn(sd, f"{D}.make_pred_c{s}.2", p, training)
        outs.append(F.conv2d(p, sd[f"{D}.make_pred_c{s}.3.weight"], sd[f"{D}.make_pred_c{s}.3.bias"], padding=1))
        ups.append(F.interpolate(x, size=size1, mode="bilinear", align_corners=False) if s > 1 else x)
        prev = x
    c = F.conv2d(torch.cat(ups, dim=1), sd[f"{D}.linear_fuse.0.weight"], sd[f"{D}.linear_fuse.0.bias"])
    c = _bn(sd, f"{D}.linear_fuse.1", c, training)
    for up, res in (("convd2x", "dense_2x"), ("convd1x", "dense_1x")):
        c = F.conv_transpose2d(c, sd[f"{D}.{up}.conv2d.weight"], sd[f"{D}.{up}.conv2d.bias"], stride=2, padding=1)
        r = F.relu(F.conv2d(c, sd[f"{D}.{res}.0.conv1.conv2d.weight"], sd[f"{D}.{res}.0.conv1.conv2d.bias"], padding=1))
        c = F.conv2d(r, sd[f"{D}.{res}.0.conv2.conv2d.weight"], sd[f"{D}.{res}.0.conv2.conv2d.bias"], padding=1) * 0.1 + c
    outs.append(F.conv2d(c, sd[f"{D}.change_probability.conv2d.weight"], sd[f"{D}.change_probability.conv2d.bias"], padding=1))
    return [torch.sigmoid(o) for o in outs] if decoder_softmax else outs


def changeformer_forward(sd, x1, x2, training=True, decoder_softmax=True, stoch: "Stoch" = None) -> List[torch.Tensor]:
    if stoch is None:
        return decoder(sd, encoder(sd, x1), encoder(sd, x2), training, decoder_softmax)
    n = x1.shape[0]                  # the encoder has no cross-sample coupling: cat(x1, x2) through it == two separate passes
    f = encoder(sd, torch.cat((x1, x2), 0), st=stoch)
    return decoder(sd, [t[:n] for t in f], [t[n:] for t in f], training, decoder_softmax)


def train_step(sd, x1, x2, mask, class_weights=(1.0, 1.0, 1.0), decoder_softmax=True, stoch=None):
    """Reference training step without multi_scale_train: the criterion sees only the last output (change_detection_trainer.py:
    166-170), so the make_pred_c* heads receive no gradient (returned as zeros)."""
    from .snunet_oracle import ce_dice_torch
    names = [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work = dict(sd)
    work.update(leaves)
    outs = changeformer_forward(work, x1, x2, True, decoder_softmax, stoch)
    loss = ce_dice_torch(outs[-1], mask, class_weights)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return loss.detach(), [o.detach() for o in outs], {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}
