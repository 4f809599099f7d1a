"""Drop-in for the reference's `models/vision_transformer.py` (ViT, Transformer, Attention, FeedForward) and for
`models/model_utilities.py:FinetunerSegmentation` (the FloodViT segmentation model: ViT encoder + head).

Same constructor keywords `ViT(image_size=, patch_size=, num_classes=, dim=, depth=, heads=, mlp_dim=, pool=, channels=,
dim_head=, dropout=, emb_dropout=)`, same registration order and state-dict keys (`pos_embedding`,
`cls_token`, `to_patch_embedding.{1,2,3}`, `transformer.norm`, `transformer.layers.{l}.0.{norm,to_qkv,to_out.0}`,
`transformer.layers.{l}.1.net.{0,1,4}`, `mlp_head`), same calls: `ViT(img) -> [B, N, dim]` tokens without the cls token
(vision_transformer.py:151-152) and `FinetunerSegmentation(encoder, configs)(img) -> [B, num_classes, 224, 224]` logits.
The sub-modules are parameter containers; the arithmetic runs in the sm_100a kernels behind `ViTSegEngine`
(vit_engine.py).  There is no eager/CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .vit_engine import ViTSegEngine


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner_dim = dim_head * heads
        if heads == 1 and dim_head == dim:
            raise NotImplementedError("project_out=False (heads == 1 and dim_head == dim) is not on the fused path")
        self.heads, self.scale = heads, dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout),
                                              FeedForward(dim, mlp_dim, dropout=dropout)]))


class _EngineHost(nn.Module):
    """Shared engine management of the trainable wrappers."""

    precision = "bf16"

    def _storage_dtype(self) -> torch.dtype:
        if self.precision == "bf16":
            return torch.bfloat16
        if self.precision == "fp32":
            return torch.float32
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision}")

    def set_ops(self, ops):
        self._ops = ops
        self._engines = {}

    def _get_ops(self, x):
        if self._ops is None:
            if not x.is_cuda:
                raise RuntimeError(f"kurosiwo_b200.{type(self).__name__} runs on a CUDA device only (no CPU fallback)")
            from .lib import default_ops
            self._ops = default_ops()
        return self._ops

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engines"] = {}
        d["_ops"] = None
        return d


class ViT(_EngineHost):
    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, pool="cls", channels=3, dim_head=64,
                 dropout=0.0, emb_dropout=0.0, precision="bf16"):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, "Image dimensions must be divisible by the patch size."
        if dropout != 0.0 or emb_dropout != 0.0:
            raise NotImplementedError("dropout > 0 is not on the fused path (the reference configs use 0: vision_transformer.py:104-105)")
        if (patch_height, patch_width) != (16, 16) or image_height != image_width:
            raise NotImplementedError("the fused patchify kernel is built for square images and 16x16 patches")
        num_patches = (image_height // patch_height) * (image_width // patch_width)
        patch_dim = channels * patch_height * patch_width
        assert pool in {"cls", "mean"}, "pool type must be either cls (cls token) or mean (mean pooling)"
        self.cfg = dict(image_size=image_height, patch_size=patch_height, channels=channels, dim=dim, depth=depth, heads=heads,
                        dim_head=dim_head, mlp_dim=mlp_dim)
        self.precision = precision
        # index 0 stands for einops' Rearrange (no parameters), so that the LayerNorm/Linear/LayerNorm keep indices 1, 2, 3
        self.to_patch_embedding = nn.Sequential(nn.Identity(), nn.LayerNorm(patch_dim), nn.Linear(patch_dim, dim), nn.LayerNorm(dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Linear(dim, num_classes)
        self._engines, self._ops = {}, None

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        """Encoder-only inference: tokens without the cls token, [B, N, dim] fp32 (vision_transformer.py:151-152).
        Training goes through FinetunerSegmentation (the trainable unit of the reference's finetune path)."""
        if self.pool == "mean":
            raise NotImplementedError("pool='mean' classification head is outside the segmentation hot path")
        wrap = getattr(self, "_wrap", None)
        if wrap is None:
            wrap = FinetunerSegmentation.__new__(FinetunerSegmentation)
            nn.Module.__init__(wrap)
            wrap.configs, wrap.pool, wrap.precision = {"num_classes": 3, "finetuning_patch_size": 16}, False, self.precision
            wrap.model = self
            wrap.head = nn.Conv2d(self.cfg["dim"], 3, kernel_size=1).to(self.pos_embedding.device)
            wrap._engines, wrap._ops = {}, self._ops
            object.__setattr__(self, "_wrap", wrap)
        with torch.no_grad():
            eng = wrap._engine_for(img)
            eng.forward(img, training=False)
            return eng.tokens().clone()


class _FinetuneFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, img, *params):
        eng = model._engine_for(img)
        ctx.engine = eng
        return eng.forward(img, training=model.training).detach().clone()

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.engine
        d = dout.contiguous()
        if d.dtype != torch.float32:
            d = d.float()
        eng.backward(d)
        flat = eng.params.grad.clone()
        grads = [flat[off:off + shape.numel()].view(shape) for off, shape in (eng.params.offsets[n] for n in eng.params.names)]
        return (None, None, *grads)


class Decoder(nn.Module):
    """models/model_utilities.py:21-48 (parameters only: the arithmetic runs in mlp_head_engine.ViTDecoderHeadEngine)."""

    def __init__(self, input_size, output_channels):
        super().__init__()
        self.deconv1 = nn.ConvTranspose2d(1024, 128, kernel_size=4, stride=2, padding=1)
        self.relu = nn.ReLU()
        self.up = nn.Upsample(scale_factor=2)
        self.deconv2 = nn.ConvTranspose2d(128, 64, kernel_size=4, stride=2, padding=1)
        self.deconv3 = nn.ConvTranspose2d(64, output_channels, kernel_size=4, stride=2, padding=1)


class FinetunerSegmentation(_EngineHost):
    """models/model_utilities.py:51-94.  configs keys: mlp, decoder, num_classes, finetuning_patch_size (as the reference)."""

    def __init__(self, encoder: ViT, configs=None, pool=False, precision=None):
        super().__init__()
        self.configs = configs
        self.model = encoder
        self.model.pool = pool
        self.pool = pool
        if pool:
            raise NotImplementedError("pool=True (one Linear over the cls token) is outside the fused path")
        if configs.get("mlp"):                                    # model_utilities.py:60-65 (mlp wins over decoder, as in the reference)
            self.head = nn.Sequential(nn.Conv2d(encoder.mlp_head.in_features, 512, kernel_size=1), nn.ReLU(),
                                      nn.Conv2d(512, configs["num_classes"], kernel_size=1))
        elif configs.get("decoder"):                              # model_utilities.py:66-69
            self.head = Decoder(encoder.mlp_head.in_features, configs["num_classes"])
        else:
            self.head = nn.Conv2d(encoder.mlp_head.in_features, configs["num_classes"], kernel_size=1)
        self.model.mlp_head = nn.Identity()
        self.precision = precision or encoder.precision
        self._engines, self._ops = {}, None

    def _engine_for(self, x: torch.Tensor) -> ViTSegEngine:
        ops = self._get_ops(x)
        key = (x.shape[0], x.shape[2], x.shape[3], self._storage_dtype(), str(x.device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines = {}
            if self.configs.get("finetuning_patch_size", 16) != 16:
                raise NotImplementedError("finetuning_patch_size must equal the encoder's 16x16 patches")
            if self.configs.get("decoder") and not self.configs.get("mlp"):
                from .mlp_head_engine import ViTDecoderHeadEngine
                eng = ViTDecoderHeadEngine(ops, self, "model.", self.model.cfg, self.configs["num_classes"], x.shape[0], x.shape[2], x.shape[3],
                                           self._storage_dtype(), x.device)
            elif self.configs.get("mlp"):
                from .mlp_head_engine import ViTMlpHeadEngine
                eng = ViTMlpHeadEngine(ops, self, "model.", self.model.cfg, self.configs["num_classes"], x.shape[0], x.shape[2], x.shape[3],
                                       self._storage_dtype(), x.device)
            else:
                eng = ViTSegEngine(ops, self, "model.", self.model.cfg, "linear", self.configs["num_classes"], x.shape[0], x.shape[2],
                                   x.shape[3], self._storage_dtype(), x.device)
            self._engines[key] = eng
        return eng

    def engine(self, x):
        return self._engine_for(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        cfg = self.model.cfg
        if x.dim() != 4 or x.shape[1] != cfg["channels"] or x.shape[2] != cfg["image_size"] or x.shape[3] != cfg["image_size"]:
            raise ValueError(f"expected [B,{cfg['channels']},{cfg['image_size']},{cfg['image_size']}], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            eng = self._engine_for(x)
            eng.params.ensure(x.device)
            return _FinetuneFunction.apply(self, x, *[p for _, p in self.named_parameters()])
        eng = self._engine_for(x)
        return eng.forward(x, training=self.training).detach().clone()


# ---------------------------------------------------------------------------------------------------------------------
# FloodViT + UPerNet head (BASELINE.json configs[3]).  Head modules mirror HF transformers' UperNetHead state-dict keys.
# ---------------------------------------------------------------------------------------------------------------------
class UperNetConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=padding, bias=False)
        self.batch_norm = nn.BatchNorm2d(out_channels)
        self.activation = nn.ReLU()


class UperNetPyramidPoolingBlock(nn.Module):
    def __init__(self, pool_scale, in_channels, channels):
        super().__init__()
        self.add_module("0", nn.AdaptiveAvgPool2d(pool_scale))
        self.add_module("1", UperNetConvModule(in_channels, channels, kernel_size=1))


class UperNetPyramidPoolingModule(nn.Module):
    def __init__(self, pool_scales, in_channels, channels):
        super().__init__()
        for i, s in enumerate(pool_scales):
            self.add_module(str(i), UperNetPyramidPoolingBlock(s, in_channels, channels))


class UperNetHead(nn.Module):
    """Parameter container with the key names / shapes of transformers.models.upernet.modeling_upernet.UperNetHead."""

    def __init__(self, in_channels, hidden_size=512, num_labels=3, pool_scales=(1, 2, 3, 6)):
        super().__init__()
        if tuple(pool_scales) != (1, 2, 3, 6):
            raise NotImplementedError("the fused head is built for pool_scales (1, 2, 3, 6) (UperNetConfig default)")
        self.classifier = nn.Conv2d(hidden_size, num_labels, kernel_size=1)
        self.psp_modules = UperNetPyramidPoolingModule(pool_scales, in_channels[-1], hidden_size)
        self.bottleneck = UperNetConvModule(in_channels[-1] + len(pool_scales) * hidden_size, hidden_size, kernel_size=3, padding=1)
        self.lateral_convs = nn.ModuleList([UperNetConvModule(c, hidden_size, kernel_size=1) for c in in_channels[:-1]])
        self.fpn_convs = nn.ModuleList([UperNetConvModule(hidden_size, hidden_size, kernel_size=3, padding=1) for _ in in_channels[:-1]])
        self.fpn_bottleneck = UperNetConvModule(len(in_channels) * hidden_size, hidden_size, kernel_size=3, padding=1)


class FloodViTUperNet(_EngineHost):
    """ViT encoder (`model`) + UPerNet head (`decode_head`): model(img[B,6,224,224]) -> logits [B, num_classes, 224, 224]."""

    def __init__(self, encoder: ViT, num_classes=3, hidden_size=512, out_indices=None, precision=None):
        super().__init__()
        self.model = encoder
        self.model.mlp_head = nn.Identity()
        D = encoder.cfg["dim"]
        self.decode_head = UperNetHead([D] * 4, hidden_size, num_classes)
        self.num_classes, self.hidden_size, self.out_indices = num_classes, hidden_size, out_indices
        self.precision = precision or encoder.precision
        self._engines, self._ops = {}, None

    def _engine_for(self, x):
        from .upernet_engine import ViTUperNetEngine
        ops = self._get_ops(x)
        key = (x.shape[0], x.shape[2], x.shape[3], self._storage_dtype(), str(x.device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines = {}
            eng = ViTUperNetEngine(ops, self, "model.", self.model.cfg, self.num_classes, x.shape[0], x.shape[2], x.shape[3],
                                   self._storage_dtype(), x.device, self.out_indices, self.hidden_size)
            self._engines[key] = eng
        return eng

    def engine(self, x):
        return self._engine_for(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        cfg = self.model.cfg
        if x.dim() != 4 or x.shape[1] != cfg["channels"] or x.shape[2] != cfg["image_size"] or x.shape[3] != cfg["image_size"]:
            raise ValueError(f"expected [B,{cfg['channels']},{cfg['image_size']},{cfg['image_size']}], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            eng = self._engine_for(x)
            eng.params.ensure(x.device)
            return _FinetuneFunction.apply(self, x, *[p for _, p in self.named_parameters()])
        eng = self._engine_for(x)
        return eng.forward(x, training=self.training).detach().clone()
