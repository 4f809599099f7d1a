"""Host-side mirror of the parts of the reference's `utilities/utilities.py` that the training hot
path consumes: loss factory, class weights / device derivation, metrics, LR scheduler factory.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .bce_and_dice import BCEandDiceLoss, FusedCrossEntropyLoss

RANDOM_EVENTS_CLASS_WEIGHTS = [0.3715753140309927, 14.009780283125977, 8.20405370357821]  # utilities.py:393-397


def update_config(config, args=None):
    """Derivations of utilities/utilities.py:350-412 that the hot path reads (class weights, device, num_channels)."""
    if config.get("weighted") and config.get("track") == "RandomEvents":
        config["class_weights"] = list(RANDOM_EVENTS_CLASS_WEIGHTS)
    else:
        config["class_weights"] = [1.0, 1.0, 1.0]
    config["device"] = f'cuda:{config["gpu"]}' if config.get("gpu") is not None else "cpu"
    if "num_channels" not in config:
        nch = len(config.get("channels", ["vv", "vh"]))
        if config.get("task") == "cd":
            config["num_channels"] = nch + (1 if config.get("dem") else 0)            # utilities.py:377-380
        else:
            config["num_channels"] = nch * len(config.get("inputs", [])) + (1 if config.get("dem") else 0)
    return config


def create_loss(configs, mode="val"):
    """utilities/utilities.py:307-347.  'ce+dice' and 'cross_entropy' both run on the fused sm_100a loss kernel."""
    weights = configs.get("class_weights", [1.0, 1.0, 1.0])
    if configs["loss_function"] == "ce+dice":
        return BCEandDiceLoss(weights=torch.tensor(weights), ignore_index=3, use_softmax=True).to(configs["device"])
    if configs["loss_function"] == "cross_entropy":           # class weights only in train mode, as the reference (:316-321)
        w = torch.tensor(weights) if mode == "train" else None
        if str(configs["device"]).startswith("cuda"):
            return FusedCrossEntropyLoss(weight=w, ignore_index=3, num_classes=configs.get("num_classes", 3)).to(configs["device"])
        return nn.CrossEntropyLoss(weight=w, ignore_index=3).to(configs["device"])
    raise NotImplementedError(f'loss_function {configs["loss_function"]} is outside the B200 hot path (SURVEY.md §8)')


class ConfusionMetrics:
    """Replacement for the five torchmetrics objects of utilities.py:228-265 (multiclass, num_classes+1=4,
    ignore_index=3, average='none', global): everything derives from one 4x4 confusion matrix kept on device."""

    def __init__(self, num_classes: int = 3, ignore_index: int = 3, device="cpu"):
        self.K, self.ignore = num_classes + 1, ignore_index
        self.mat = torch.zeros(self.K, self.K, dtype=torch.int64, device=device)

    def reset(self):
        self.mat.zero_()

    def update(self, preds: torch.Tensor, target: torch.Tensor):
        if preds.is_cuda and preds.dtype == torch.uint8 and target.dtype == torch.int64 and self.K == 4 and preds.is_contiguous() \
                and target.is_contiguous():
            from .lib import default_ops        # one kernel, no host sync (the uint8 argmax map comes from the loss kernel)
            default_ops().confusion_update(preds, target, self.K, self.ignore, self.mat)
            return
        t = target.reshape(-1)
        p = preds.reshape(-1).to(torch.int64)
        keep = t != self.ignore
        idx = t[keep] * self.K + p[keep]
        self.mat += torch.bincount(idx, minlength=self.K * self.K).view(self.K, self.K)

    def compute(self):
        m = self.mat.double()
        tp = m.diag()
        fn = m.sum(1) - tp
        fp = m.sum(0) - tp
        def safe(a, b):
            return torch.where(b > 0, a / b.clamp_min(1), torch.zeros_like(a))
        acc = safe(tp, tp + fn)
        prec = safe(tp, tp + fp)
        rec = acc.clone()
        f1 = safe(2 * tp, 2 * tp + fp + fn)
        iou = safe(tp, tp + fp + fn)
        return acc, f1, prec, rec, iou


def initialize_metrics(configs, mode="all"):
    return ConfusionMetrics(configs["num_classes"], 3, configs["device"])


def init_lr_scheduler(optimizer, configs, model_configs, model_name=None, steps=None):
    """utilities/utilities.py:268-304 (per-epoch schedulers)."""
    sched = model_configs[model_name]["lr_schedule"] if model_name is not None else model_configs.get("lr_schedule")
    if sched == "cosine":
        return torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, steps)
    if sched is None:
        return torch.optim.lr_scheduler.LambdaLR(optimizer, lambda _: 1, last_epoch=-1)
    if sched == "linear":
        return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda e: 1.0 - e / float(configs["epochs"] + 1))
    raise NotImplementedError(f"{sched} LR scheduling is not yet implemented!")
