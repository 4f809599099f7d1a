"""Drop-in for the reference's `utilities/bce_and_dice.py` + `utilities/dice.py`.

`BCEandDiceLoss(weights, ignore_index, use_softmax)` / `forward(preds[N,C,H,W], lbl[N,H,W] int64)`
-> 0-dim loss supporting `.item()` and `.backward()`; same argument checks and exception types as
utilities/dice.py:97-109.  Forward value, gradient and the argmax class map come from ONE fused
sm_100a kernel (csrc/loss.cu); there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class _CeDiceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, lbl, weights, ignore_index, state, dice_weight=1.0):
        from .lib import default_ops
        ops = state.get("ops") or default_ops()
        logits = preds.detach()
        if logits.dtype != torch.float32:
            logits = logits.float()  # CE/softmax run in fp32 under autocast as well
        logits = logits.contiguous()
        N = logits.shape[0]
        need_grad = preds.requires_grad
        key = (tuple(logits.shape), str(logits.device))
        if state.get("key") != key:
            state["key"] = key
            state["ws"] = ops.ce_dice_workspace(N, logits.device)
            state["loss3"] = torch.zeros(3, dtype=torch.float32, device=logits.device)
            state["dlogits"] = torch.empty_like(logits)
            state["pred"] = torch.empty((N,) + tuple(logits.shape[2:]), dtype=torch.uint8, device=logits.device)
        ops.ce_dice(logits, lbl.contiguous(), weights, ignore_index, 1.0, state["loss3"],
                    state["dlogits"] if need_grad else None, state["pred"], state["ws"], dice_weight)
        ctx.dlogits = state["dlogits"] if need_grad else None
        ctx.in_dtype = preds.dtype
        return state["loss3"][0].clone()

    @staticmethod
    def backward(ctx, gout):
        g = ctx.dlogits * gout
        if g.dtype != ctx.in_dtype:
            g = g.to(ctx.in_dtype)
        return g, None, None, None, None, None


class DiceLoss(nn.Module):
    """Kept for API parity (utilities/dice.py:62); the fused kernel always evaluates CE+Dice together."""

    def __init__(self, ignore_index=None, use_softmax=False) -> None:
        super().__init__()
        self.eps = 1e-6
        self.ignore_index = ignore_index
        self.use_softmax = use_softmax


class BCEandDiceLoss(nn.Module):
    dice_weight = 1.0

    def __init__(self, weights=None, ignore_index=None, use_softmax=False):
        super().__init__()
        if not use_softmax:
            raise NotImplementedError("the reference only builds this loss with use_softmax=True (utilities/utilities.py:347)")
        w = torch.as_tensor(weights, dtype=torch.float32).clone()
        self.register_buffer("weight", w)
        # ignore_index=None in the reference means "no pixel is ignored"; -100 is torch's own default
        self.ignore_index = -100 if ignore_index is None else int(ignore_index)
        self.dice = DiceLoss(ignore_index=ignore_index, use_softmax=use_softmax)
        self._state = {}

    @property
    def last_pred(self) -> Optional[torch.Tensor]:
        """uint8 argmax class map of the last forward (== preds.argmax(1)), emitted by the same kernel."""
        return self._state.get("pred")

    @property
    def last_parts(self) -> Optional[torch.Tensor]:
        """fp32 [3] = (total, dice, ce) of the last forward."""
        return self._state.get("loss3")

    def forward(self, preds: torch.Tensor, lbl: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(preds):
            raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(preds)))
        if not len(preds.shape) == 4:
            raise ValueError("Invalid input shape, we expect BxNxHxW. Got: {}".format(preds.shape))
        if not preds.shape[-2:] == lbl.shape[-2:]:
            raise ValueError("input and target shapes must be the same. Got: {}".format(preds.shape))
        if not preds.device == lbl.device:
            raise ValueError("input and target must be in the same device. Got: {}".format(preds.device, lbl.device))
        if not len(lbl.shape) == 3:
            raise ValueError("Invalid depth shape, we expect BxHxW. Got: {}".format(lbl.shape))
        if not lbl.dtype == torch.int64:
            raise ValueError("labels must be of the same dtype torch.int64. Got: {}".format(lbl.dtype))
        if not preds.is_cuda:
            raise RuntimeError("kurosiwo_b200.BCEandDiceLoss runs on a CUDA device only (no CPU fallback)")
        if self.weight.device != preds.device:
            self.weight = self.weight.to(preds.device)
        return _CeDiceFunction.apply(preds, lbl, self.weight, self.ignore_index, self._state, float(self.dice_weight))

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_state"] = {}
        return d


class FusedCrossEntropyLoss(BCEandDiceLoss):
    """nn.CrossEntropyLoss(weight=w, ignore_index=i) - the reference's default criterion (utilities/utilities.py:308-321) - on the
    same fused kernel with the Dice term switched off (dice_weight = 0): value, gradient and the argmax map in one pass."""
    dice_weight = 0.0

    def __init__(self, weight=None, ignore_index=-100, num_classes=3):
        w = torch.ones(num_classes) if weight is None else weight
        super().__init__(weights=w, ignore_index=ignore_index, use_softmax=True)
