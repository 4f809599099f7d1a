"""Drop-in for the reference's `models/snunet.py` (SNUNet_ECAM, conv_block_nested, up, ChannelAttention).

Same constructor signature `SNUNet_ECAM(in_channels, out_ch, base_channel=32)`, same state-dict keys,
shapes and initialisation (reference models/snunet.py:65-115), same call `model(xA, xB)` returning
`[B, out_ch, H, W]` float32 logits that support `.argmax(1)` and `.backward()`.  The sub-modules
are parameter containers only: the arithmetic runs in the sm_100a kernels behind
`SNUNetEngine` (snunet_engine.py).  There is no eager/CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .snunet_engine import DEC_ORDER, SNUNetEngine


class conv_block_nested(nn.Module):
    """Parameters of conv3x3 -> BN -> ReLU -> conv3x3 -> BN -> (+conv1 output) -> ReLU (snunet.py:11-29)."""

    def __init__(self, in_ch: int, mid_ch: int, out_ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, mid_ch, 3, padding=1, bias=True)
        self.bn1 = nn.BatchNorm2d(mid_ch)
        self.conv2 = nn.Conv2d(mid_ch, out_ch, 3, padding=1, bias=True)
        self.bn2 = nn.BatchNorm2d(out_ch)

    def forward(self, x):  # pragma: no cover - containers are never called
        raise RuntimeError("conv_block_nested is executed by SNUNetEngine, not eagerly")


class up(nn.Module):
    """Parameters of ConvTranspose2d(C, C, 2, stride=2) (snunet.py:32-46)."""

    def __init__(self, in_ch: int, bilinear: bool = False):
        super().__init__()
        if bilinear:
            raise NotImplementedError("the reference only instantiates the transposed-conv variant (snunet.py:77-101)")
        self.up = nn.ConvTranspose2d(in_ch, in_ch, 2, stride=2)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("up is executed by SNUNetEngine, not eagerly")


class ChannelAttention(nn.Module):
    """Parameters of the two bias-free 1x1 FCs of the channel attention (snunet.py:49-62)."""

    def __init__(self, in_channels: int, ratio: int = 16):
        super().__init__()
        self.fc1 = nn.Conv2d(in_channels, in_channels // ratio, 1, bias=False)
        self.fc2 = nn.Conv2d(in_channels // ratio, in_channels, 1, bias=False)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("ChannelAttention is executed by SNUNetEngine, not eagerly")


class _SNUNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, xA, xB, *params):
        eng = model._engine_for(xA)
        ctx.engine = eng
        # a private copy: the engine's logits buffer is overwritten by the next forward
        return eng.forward(xA, xB, training=model.training).detach().clone()

    @staticmethod
    def backward(ctx, dlogits):
        eng = ctx.engine
        dl = dlogits.contiguous()
        if dl.dtype != torch.float32:
            dl = dl.float()
        eng.backward(dl)
        # hand autograd a private copy: AccumulateGrad may keep (or add into) these tensors, and the
        # engine's flat gradient buffer is overwritten by the next backward.
        flat = eng.params.grad.clone()
        grads = [flat[off:off + shape.numel()].view(shape) for off, shape in (eng.params.offsets[n] for n in eng.params.names)]
        return (None, None, None, *grads)


class SNUNet_ECAM(nn.Module):
    """SNUNet-CD with ECAM.  `precision`: "bf16" (perf mode, tcgen05) or "fp32" (parity mode)."""

    def __init__(self, in_channels: int, out_ch: int, base_channel: int = 32, precision: str = "bf16"):
        super().__init__()
        n1 = base_channel
        f = [n1, n1 * 2, n1 * 4, n1 * 8, n1 * 16]
        self.in_channels, self.out_ch, self.base_channel = in_channels, out_ch, base_channel
        self.precision = precision
        # registration order follows the reference constructor so that state_dict()/parameters() order matches
        for l in range(5):
            cin = in_channels if l == 0 else f[l - 1]
            setattr(self, f"conv{l}_0", conv_block_nested(cin, f[l], f[l]))
            if l >= 1:
                setattr(self, f"Up{l}_0", up(f[l]))
        for j in range(1, 5):
            for l in range(0, 5 - j):
                setattr(self, f"conv{l}_{j}", conv_block_nested(f[l] * (j + 1) + f[l + 1], f[l], f[l]))
                if l >= 1 and j <= 3 and (l - 1, j + 1) in DEC_ORDER:
                    setattr(self, f"Up{l}_{j}", up(f[l]))
        self.ca = ChannelAttention(f[0] * 4, ratio=16)
        self.ca1 = ChannelAttention(f[0], ratio=16 // 4)
        self.conv_final = nn.Conv2d(f[0] * 4, out_ch, kernel_size=1)
        for m in self.modules():  # snunet.py:110-115
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._engines = {}
        self._ops = None

    # -- engine management ---------------------------------------------------------------------
    def _storage_dtype(self) -> torch.dtype:
        if self.precision == "bf16":
            return torch.bfloat16
        if self.precision == "fp32":
            return torch.float32
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision}")

    def set_ops(self, ops):
        """Inject the op backend (tests use this to check the host schedule); default = CUDA library."""
        self._ops = ops
        self._engines = {}

    def _engine_for(self, x: torch.Tensor) -> SNUNetEngine:
        if self._ops is None:
            if not x.is_cuda:
                raise RuntimeError("kurosiwo_b200.SNUNet_ECAM runs on a CUDA device only (no CPU fallback)")
            from .lib import default_ops
            self._ops = default_ops()
        key = (x.shape[0], x.shape[2], x.shape[3], self._storage_dtype(), str(x.device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines = {}  # one plan at a time: activations are sized for the batch
            eng = SNUNetEngine(self._ops, self, self.in_channels, self.out_ch, self.base_channel,
                               x.shape[0], x.shape[2], x.shape[3], self._storage_dtype(), x.device)
            self._engines[key] = eng
        return eng

    def engine(self, x: torch.Tensor) -> SNUNetEngine:
        return self._engine_for(x)

    def forward(self, xA: torch.Tensor, xB: torch.Tensor) -> torch.Tensor:
        if xA.shape != xB.shape or xA.dim() != 4 or xA.shape[1] != self.in_channels:
            raise ValueError(f"expected two [B,{self.in_channels},H,W] tensors, got {tuple(xA.shape)} and {tuple(xB.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            eng = self._engine_for(xA)
            eng.params.ensure(xA.device)
            return _SNUNetFunction.apply(self, xA, xB, *[p for _, p in self.named_parameters()])
        eng = self._engine_for(xA)
        return eng.forward(xA, xB, training=self.training).detach().clone()   # the fused trainer reads engine.logits directly

    def __getstate__(self):  # torch.save(model) must not pickle device plans (segmentation_trainer.py:255 style)
        d = dict(self.__dict__)
        d["_engines"] = {}
        d["_ops"] = None
        return d
