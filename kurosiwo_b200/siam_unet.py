"""Drop-in for the reference's `models/siam_conc.py` (SiamUnet_conc) and `models/siam_diff.py` (SiamUnet_diff).

Same constructor `SiamUnet_conc(input_nbr, label_nbr)`, same registration order, state-dict keys and shapes (reference
models/siam_conc.py:13-93: conv11..conv43 nn.Conv2d, upconv1..4 / conv43d..conv11d nn.ConvTranspose2d, bnXX nn.BatchNorm2d,
doXX nn.Dropout2d), same call `model(x1, x2)` returning `[B, label_nbr, H, W]` float32 probabilities (conc, Softmax) or
log-probabilities (diff, LogSoftmax) that support `.argmax(1)` and `.backward()`.  The sub-modules are parameter containers:
the arithmetic runs in the sm_100a kernels behind `SiamUNetEngine` (siam_engine.py).  There is no eager/CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .siam_engine import DEC, ENC, UP_BEFORE, SiamUNetEngine


class _SiamFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x1, x2, *params):
        eng = model._engine_for(x1)
        ctx.engine = eng
        return eng.forward(x1, x2, training=model.training).detach().clone()

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.engine
        d = dout.contiguous()
        if d.dtype != torch.float32:
            d = d.float()
        eng.backward(d)
        flat = eng.params.grad.clone()
        grads = [flat[off:off + shape.numel()].view(shape) for off, shape in (eng.params.offsets[n] for n in eng.params.names)]
        return (None, None, None, *grads)


class _SiamUnet(nn.Module):
    kind = "conc"

    def __init__(self, input_nbr: int, label_nbr: int, precision: str = "bf16"):
        super().__init__()
        self.input_nbr, self.label_nbr, self.precision = input_nbr, label_nbr, precision
        self.dropout_p = 0.2
        cin = input_nbr
        for n, c in ENC:                                   # siam_conc.py:19-50
            setattr(self, f"conv{n}", nn.Conv2d(cin, c, kernel_size=3, padding=1))
            setattr(self, f"bn{n}", nn.BatchNorm2d(c))
            setattr(self, f"do{n}", nn.Dropout2d(p=0.2))
            cin = c
        mult = 3 if self.kind == "conc" else 2
        prev = 128
        for n, c, _ in DEC:                                # siam_conc.py:53-90
            if n in UP_BEFORE:
                setattr(self, UP_BEFORE[n], nn.ConvTranspose2d(prev, prev, kernel_size=3, padding=1, stride=2, output_padding=1))
                setattr(self, f"conv{n}", nn.ConvTranspose2d(prev * mult, c, kernel_size=3, padding=1))
            else:
                setattr(self, f"conv{n}", nn.ConvTranspose2d(prev, c, kernel_size=3, padding=1))
            setattr(self, f"bn{n}", nn.BatchNorm2d(c))
            setattr(self, f"do{n}", nn.Dropout2d(p=0.2))
            prev = c
        self.conv11d = nn.ConvTranspose2d(16, label_nbr, kernel_size=3, padding=1)
        self.sm = nn.Softmax(dim=1) if self.kind == "conc" else nn.LogSoftmax(dim=1)
        self._engines = {}
        self._ops = None

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

    def _engine_for(self, x: torch.Tensor) -> SiamUNetEngine:
        if self._ops is None:
            if not x.is_cuda:
                raise RuntimeError(f"kurosiwo_b200.{type(self).__name__} runs on a CUDA device only (no CPU fallback)")
            from .lib import default_ops
            self._ops = default_ops()
        key = (x.shape[0], x.shape[2], x.shape[3], self._storage_dtype(), str(x.device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines = {}
            eng = SiamUNetEngine(self._ops, self, self.kind, self.input_nbr, self.label_nbr, x.shape[0], x.shape[2], x.shape[3],
                                 self._storage_dtype(), x.device)
            self._engines[key] = eng
        return eng

    def engine(self, x: torch.Tensor) -> SiamUNetEngine:
        return self._engine_for(x)

    def forward(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        if x1.shape != x2.shape or x1.dim() != 4 or x1.shape[1] != self.input_nbr:
            raise ValueError(f"expected two [B,{self.input_nbr},H,W] tensors, got {tuple(x1.shape)} and {tuple(x2.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            eng = self._engine_for(x1)
            eng.params.ensure(x1.device)
            return _SiamFunction.apply(self, x1, x2, *[p for _, p in self.named_parameters()])
        eng = self._engine_for(x1)
        return eng.forward(x1, x2, training=self.training).detach().clone()

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engines"] = {}
        d["_ops"] = None
        return d


class SiamUnet_conc(_SiamUnet):
    """FC-Siam-conc (models/siam_conc.py:13)."""
    kind = "conc"


class SiamUnet_diff(_SiamUnet):
    """FC-Siam-diff (models/siam_diff.py:13)."""
    kind = "diff"
