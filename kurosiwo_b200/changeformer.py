"""Drop-in for the reference's `models/changeformer.py` (ChangeFormerV6 and its sub-modules).

Same constructor `ChangeFormerV6(input_nc, output_nc, decoder_softmax, embed_dim)`, same registration order and the 373 state-dict
keys / shapes of the reference (`Tenc_x2.patch_embed{s}.{proj,norm}`, `Tenc_x2.block{s}.{i}.{norm1,attn.{q,kv,proj,
sr,norm},norm2,mlp.{fc1,dwconv.dwconv,fc2}}`, `Tenc_x2.norm{s}`, `TDec_x2.{linear_c*,diff_c*,make_pred_c*,linear_fuse,convd2x,dense_2x,
convd1x,dense_1x,change_probability}`), same initialisation rules (changeformer.py:115-124, 408-417) and the same call:
`model(x1, x2)` returns the LIST [p_c4, p_c3, p_c2, p_c1, cp] of post-Sigmoid maps whose last element the trainer uses
(change_detection_trainer.py:148,166).  The sub-modules are parameter containers; the arithmetic runs in the sm_100a kernels
behind `ChangeFormerEngine` (cformer_engine.py).  There is no eager/CPU fallback.
Dropout / attention dropout / DropPath (0.1 each, :652-654) are active in train mode with a device-side stateless RNG
(`model.drop_rate`, `model.attn_drop`, `model.drop_path_rate`; set them to 0 for deterministic parity runs).
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn

from .cformer_engine import DEPTHS, EMBED_DIMS, HEADS, SR, ChangeFormerEngine


def _init_weights(m):          # changeformer.py:115-124 (identical in every sub-module)
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
        fan_out //= m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


class DWConv(nn.Module):
    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self.apply(_init_weights)


class MLP(nn.Module):
    """Linear Embedding (changeformer.py:135-146)."""

    def __init__(self, input_dim=2048, embed_dim=768):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0., sr_ratio=1):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim, self.num_heads, self.sr_ratio = dim, num_heads, sr_ratio
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.norm = nn.LayerNorm(dim)
        self.apply(_init_weights)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, sr_ratio=1):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop, sr_ratio=sr_ratio)
        self.drop_path = nn.Identity()
        self.drop_path_prob = drop_path
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), drop=drop)
        self.apply(_init_weights)


class OverlapPatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=7, stride=4, in_chans=3, embed_dim=768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride, padding=(patch_size // 2, patch_size // 2))
        self.norm = nn.LayerNorm(embed_dim)
        self.apply(_init_weights)


class EncoderTransformer_v3(nn.Module):
    def __init__(self, img_size=256, patch_size=3, in_chans=3, num_classes=2, embed_dims=(32, 64, 128, 256), num_heads=(2, 2, 4, 8),
                 mlp_ratios=(4, 4, 4, 4), qkv_bias=True, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm,
                 depths=(3, 3, 6, 18), sr_ratios=(8, 4, 2, 1)):
        super().__init__()
        self.depths, self.embed_dims = list(depths), list(embed_dims)
        self.patch_embed1 = OverlapPatchEmbed(img_size, 7, 4, in_chans, embed_dims[0])
        self.patch_embed2 = OverlapPatchEmbed(img_size // 4, patch_size, 2, embed_dims[0], embed_dims[1])
        self.patch_embed3 = OverlapPatchEmbed(img_size // 8, patch_size, 2, embed_dims[1], embed_dims[2])
        self.patch_embed4 = OverlapPatchEmbed(img_size // 16, patch_size, 2, embed_dims[2], embed_dims[3])
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for s in range(4):
            blocks = nn.ModuleList([Block(dim=embed_dims[s], num_heads=num_heads[s], mlp_ratio=mlp_ratios[s], qkv_bias=qkv_bias, drop=drop_rate,
                                          attn_drop=attn_drop_rate, drop_path=dpr[cur + i], norm_layer=norm_layer, sr_ratio=sr_ratios[s])
                                    for i in range(depths[s])])
            setattr(self, f"block{s + 1}", blocks)
            setattr(self, f"norm{s + 1}", norm_layer(embed_dims[s]))
            cur += depths[s]
        self.apply(_init_weights)


def conv_diff(in_channels, out_channels):
    return nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.ReLU(), nn.BatchNorm2d(out_channels),
                         nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.ReLU())


def make_prediction(in_channels, out_channels):
    return nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.ReLU(), nn.BatchNorm2d(out_channels),
                         nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1))


class ConvLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding):
        super().__init__()
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)


class UpsampleConvLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.conv2d = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=stride, padding=1)


class ResidualBlock(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv1 = ConvLayer(channels, channels, kernel_size=3, stride=1, padding=1)
        self.conv2 = ConvLayer(channels, channels, kernel_size=3, stride=1, padding=1)
        self.relu = nn.ReLU()


class DecoderTransformer_v3(nn.Module):
    def __init__(self, in_channels=(32, 64, 128, 256), embedding_dim=64, output_nc=2, decoder_softmax=False):
        super().__init__()
        E = embedding_dim
        self.embedding_dim, self.output_nc = E, output_nc
        c1, c2, c3, c4 = in_channels
        self.linear_c4, self.linear_c3 = MLP(c4, E), MLP(c3, E)
        self.linear_c2, self.linear_c1 = MLP(c2, E), MLP(c1, E)
        self.diff_c4, self.diff_c3, self.diff_c2, self.diff_c1 = (conv_diff(2 * E, E) for _ in range(4))
        self.make_pred_c4, self.make_pred_c3, self.make_pred_c2, self.make_pred_c1 = (make_prediction(E, output_nc) for _ in range(4))
        self.linear_fuse = nn.Sequential(nn.Conv2d(E * 4, E, kernel_size=1), nn.BatchNorm2d(E))
        self.convd2x = UpsampleConvLayer(E, E, kernel_size=4, stride=2)
        self.dense_2x = nn.Sequential(ResidualBlock(E))
        self.convd1x = UpsampleConvLayer(E, E, kernel_size=4, stride=2)
        self.dense_1x = nn.Sequential(ResidualBlock(E))
        self.change_probability = ConvLayer(E, output_nc, kernel_size=3, stride=1, padding=1)
        self.output_softmax = decoder_softmax
        self.active = nn.Sigmoid()


class _CFFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x1, x2, *params):
        eng = model._engine_for(x1)
        ctx.engine = eng
        eng.forward(x1, x2, training=model.training)
        outs = [o.detach().clone() for o in eng.outputs()]
        ctx.mark_non_differentiable(*outs[:4])          # side outputs carry no gradient on the fused path (no multi_scale_train)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        eng = ctx.engine
        d = douts[-1].contiguous()
        if d.dtype != torch.float32:
            d = d.float()
        eng.backward(d)
        flat = eng.params.grad.clone()
        grads = [flat[off:off + shape.numel()].view(shape) for off, shape in (eng.params.offsets[n] for n in eng.params.names)]
        return (None, None, None, *grads)


class ChangeFormerV6(nn.Module):
    def __init__(self, input_nc=3, output_nc=2, decoder_softmax=False, embed_dim=256, precision="bf16"):
        super().__init__()
        self.embed_dims, self.depths, self.embedding_dim = list(EMBED_DIMS), list(DEPTHS), embed_dim
        self.drop_rate, self.attn_drop, self.drop_path_rate = 0.1, 0.1, 0.1       # reference values (:652-654)
        self.input_nc, self.output_nc, self.decoder_softmax, self.precision = input_nc, output_nc, decoder_softmax, precision
        self.Tenc_x2 = EncoderTransformer_v3(img_size=256, patch_size=7, in_chans=input_nc, num_classes=output_nc, embed_dims=self.embed_dims,
                                             num_heads=HEADS, mlp_ratios=[4, 4, 4, 4], qkv_bias=True, drop_rate=self.drop_rate,
                                             attn_drop_rate=self.attn_drop, drop_path_rate=self.drop_path_rate,
                                             norm_layer=partial(nn.LayerNorm, eps=1e-6), depths=self.depths, sr_ratios=SR)
        self.TDec_x2 = DecoderTransformer_v3(in_channels=self.embed_dims, embedding_dim=embed_dim, output_nc=output_nc,
                                             decoder_softmax=decoder_softmax)
        self._engines, self._ops = {}, None

    def _storage_dtype(self):
        if self.precision == "bf16":
            return torch.bfloat16
        if self.precision == "fp32":
            return torch.float32
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision}")

    def set_ops(self, ops):
        self._ops = ops
        self._engines = {}

    def _engine_for(self, x) -> ChangeFormerEngine:
        if self._ops is None:
            if not x.is_cuda:
                raise RuntimeError("kurosiwo_b200.ChangeFormerV6 runs on a CUDA device only (no CPU fallback)")
            from .lib import default_ops
            self._ops = default_ops()
        key = (x.shape[0], x.shape[2], x.shape[3], self._storage_dtype(), str(x.device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines = {}
            eng = ChangeFormerEngine(self._ops, self, self.input_nc, self.output_nc, self.embedding_dim, self.decoder_softmax,
                                     x.shape[0], x.shape[2], x.shape[3], self._storage_dtype(), x.device)
            self._engines[key] = eng
        return eng

    def engine(self, x):
        return self._engine_for(x)

    def forward(self, x1, x2):
        if x1.shape != x2.shape or x1.dim() != 4 or x1.shape[1] != self.input_nc:
            raise ValueError(f"expected two [B,{self.input_nc},H,W] tensors, got {tuple(x1.shape)} and {tuple(x2.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            eng = self._engine_for(x1)
            eng.params.ensure(x1.device)
            return list(_CFFunction.apply(self, x1, x2, *[p for _, p in self.named_parameters()]))
        eng = self._engine_for(x1)
        eng.forward(x1, x2, training=self.training)
        return [o.detach().clone() for o in eng.outputs()]

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engines"] = {}
        d["_ops"] = None
        return d
