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


def metrics_from_confusion(mat: torch.Tensor):
    """(accuracy, f1, precision, recall, iou) per class from KxK confusion matrices [..., K, K] (rows = target, cols = prediction):
    the numbers the five torchmetrics objects of utilities.py:228-265 report (multiclass, average='none', ignore_index dropped)."""
    m = mat.double()
    tp = m.diagonal(dim1=-2, dim2=-1)
    fn = m.sum(-1) - tp
    fp = m.sum(-2) - tp

    def safe(a, b):
        return torch.where(b > 0, a / b.clamp_min(1), torch.zeros_like(a))
    acc = safe(tp, tp + fn)
    return acc, safe(2 * tp, 2 * tp + fp + fn), safe(tp, tp + fp), acc.clone(), safe(tp, tp + fp + fn)


def water_only_fscore(mat: torch.Tensor) -> torch.Tensor:
    """`evaluate_water` (change_detection_trainer.py:328,408-413,534): F1 of the 2-class problem {no water, water} obtained by
    relabelling class 2 (flood) as class 1 in predictions AND labels - i.e. the confusion matrix with rows/columns 1 and 2 merged.
    Returns the per-class F1 [2]; the reference reports index 1."""
    m = mat.double()[:3, :3]
    w = torch.stack([torch.stack([m[0, 0], m[0, 1] + m[0, 2]]), torch.stack([m[1, 0] + m[2, 0], m[1:3, 1:3].sum()])])
    return metrics_from_confusion(w)[1]


class GroupedConfusionMetrics(ConfusionMetrics):
    """The global metric set plus the per-activation (AOI) and per-climate-zone sets of change_detection_trainer.py:34-38,184-199,
    :331-338,445-472 (and segmentation_trainer.py:407-512), all fed by ONE `ks_confusion_update_grouped` launch per batch keyed by
    the batch's `activ` / `clz` vectors.  activations: the dataset's activation ids (loader.dataset.activations); zones are 1..3."""

    def __init__(self, num_classes: int = 3, ignore_index: int = 3, device="cpu", activations=None, zones: bool = False):
        super().__init__(num_classes, ignore_index, device)
        self.device = torch.device(device)
        self.activations = sorted(int(a) for a in activations) if activations is not None else None
        self.mat_aoi = torch.zeros(len(self.activations), self.K, self.K, dtype=torch.int64, device=device) if self.activations else None
        self.mat_zone = torch.zeros(3, self.K, self.K, dtype=torch.int64, device=device) if zones else None
        self.samples_per_zone = {1: 0, 2: 0, 3: 0}

    def reset(self):
        super().reset()
        for m in (self.mat_aoi, self.mat_zone):
            if m is not None:
                m.zero_()
        self.samples_per_zone = {1: 0, 2: 0, 3: 0}

    def _keys(self, activ, clz):
        ka = kb = None
        if self.mat_aoi is not None and activ is not None:
            a = torch.as_tensor(activ).reshape(-1).cpu().to(torch.int64)
            table = torch.tensor(self.activations, dtype=torch.int64)
            idx = torch.searchsorted(table, a).clamp_(max=len(self.activations) - 1)
            ka = torch.where(table[idx] == a, idx, torch.full_like(idx, -1)).to(torch.int32)
        if self.mat_zone is not None and clz is not None:
            z = torch.as_tensor(clz).reshape(-1).cpu().to(torch.int64)
            kb = (z - 1).to(torch.int32)
            for k in (1, 2, 3):
                self.samples_per_zone[k] += int((z == k).sum())
        return ka, kb

    def update(self, preds: torch.Tensor, target: torch.Tensor, activ=None, clz=None):
        ka, kb = self._keys(activ, clz)
        if ka is None and kb is None:
            return super().update(preds, target)
        if preds.is_cuda and preds.dtype == torch.uint8 and target.dtype == torch.int64 and self.K == 4 and preds.is_contiguous() \
                and target.is_contiguous():
            from .lib import default_ops
            dev = preds.device
            default_ops().confusion_update_grouped(preds, target, self.K, self.ignore, self.mat,
                                                   None if ka is None else ka.to(dev, non_blocking=True), self.mat_aoi if ka is not None else None,
                                                   None if kb is None else kb.to(dev, non_blocking=True), self.mat_zone if kb is not None else None)
            return
        for s in range(target.shape[0]):                       # CPU tensors / non-uint8 predictions: torch ops, same counts
            one = torch.zeros(self.K, self.K, dtype=torch.int64, device=self.mat.device)
            t, p = target[s].reshape(-1), preds[s].reshape(-1).to(torch.int64)
            keep = t != self.ignore
            one += torch.bincount(t[keep] * self.K + p[keep], minlength=self.K * self.K).view(self.K, self.K)
            self.mat += one
            if ka is not None and int(ka[s]) >= 0:
                self.mat_aoi[int(ka[s])] += one
            if kb is not None and 0 <= int(kb[s]) < 3:
                self.mat_zone[int(kb[s])] += one

    def compute_aoi(self):
        """{activation id: (acc, f1, prec, rec, iou)} for the activations that received samples."""
        if self.mat_aoi is None:
            return {}
        seen = self.mat_aoi.sum((1, 2)).cpu()
        out = metrics_from_confusion(self.mat_aoi)
        return {a: tuple(o[i] for o in out) for i, a in enumerate(self.activations) if int(seen[i]) > 0}

    def compute_zones(self):
        if self.mat_zone is None:
            return {}
        out = metrics_from_confusion(self.mat_zone)
        return {z: tuple(o[z - 1] for o in out) for z in (1, 2, 3) if self.samples_per_zone[z] > 0}

    def water_fscore(self):
        return water_only_fscore(self.mat)


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
