"""Mirror of the reference's `training/segmentation_trainer.py` entry points for the B200 path.

`train_semantic_segmentation(model, train_loader, val_loader, test_loader, configs, model_configs)` and
`eval_semantic_segmentation(model, loader, configs, settype=..., model_configs=...)` keep the reference signatures (:16-18,
:258-266), batch tuple layout (dataset/Dataset.py:826-839) and the input stacking of :138-144 (`image = cat(post_event,
pre_event_1, pre_event_2)` for the three-date configuration).  The inner loop (:54-170) is replaced by ONE fused engine step
(forward, CE+Dice+argmax, backward, Adam) with no per-iteration host sync and a device-side confusion matrix.
"""
from __future__ import annotations

from pathlib import Path

import torch

from .change_detection_trainer import CLASS_LABELS, LAST_EVAL, _rank, preprocess_raw, sync_buffers, unpack_batch
from .host_pipeline import HostPipelineMixin, lookahead
from .utilities import GroupedConfusionMetrics, create_loss, init_lr_scheduler
from .vision_transformer import FinetunerSegmentation, FloodViTUperNet


def stack_inputs(b, configs, device):
    """segmentation_trainer.py:106-144: post_event first, then the pre-event dates named in configs['inputs'] (+ DEM)."""
    b = dict(b)
    for k in ("post_event", "pre_event_1", "pre_event_2"):      # raw_input: clamp / nan_to_num / Normalize on the device copy
        if configs.get("raw_input") and b.get(k) is not None:
            b[k] = preprocess_raw(b[k].to(device, non_blocking=True), configs)
    image = b["post_event"].to(device, non_blocking=True)
    if configs.get("dem"):
        image = torch.cat((image, b["dem"].to(device, non_blocking=True)), dim=1)
    inputs = configs["inputs"]
    if inputs == ["post_event"]:
        return image
    s = set(inputs)
    if s == {"pre_event_1", "post_event"}:
        return torch.cat((image, b["pre_event_1"].to(device, non_blocking=True)), dim=1)
    if s == {"pre_event_2", "post_event"}:
        return torch.cat((image, b["pre_event_2"].to(device, non_blocking=True)), dim=1)
    if s == {"pre_event_1", "pre_event_2", "post_event"}:
        return torch.cat((image, b["pre_event_1"].to(device, non_blocking=True), b["pre_event_2"].to(device, non_blocking=True)), dim=1)
    print('Invalid configuration for "inputs". Exiting...')
    raise SystemExit(1)


class FusedSegStepper(HostPipelineMixin):
    """Owns the engine-side training state of a segmentation model (the public fast path)."""

    def __init__(self, model, configs, model_configs, process_group=None):
        if not isinstance(model, (FinetunerSegmentation, FloodViTUperNet)):
            raise TypeError("the fused segmentation step is implemented for kurosiwo_b200's FinetunerSegmentation / FloodViTUperNet")
        if configs.get("loss_function", "ce+dice") not in ("ce+dice", "cross_entropy"):
            raise NotImplementedError("the fused step computes CE+Dice (utilities/bce_and_dice.py) or plain cross-entropy "
                                      "(utilities/utilities.py:308-321); other losses are outside the B200 hot path")
        self.model, self.configs, self.model_configs, self.pg = model, configs, model_configs, process_group
        self.engine = None
        self.lr = float(model_configs["learning_rate"])

    def _engine(self, x):
        eng = self.model.engine(x)
        if eng is not self.engine:
            eng.init_training(class_weights=self.configs.get("class_weights", [1.0, 1.0, 1.0]), ignore_index=3, lr=self.lr,
                              betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=self.pg,
                              dice_weight=0.0 if self.configs.get("loss_function") == "cross_entropy" else 1.0)   # torch.optim.Adam(lr) (:36)
            eng.adopt_training_state(self.engine)   # another batch geometry (e.g. the ragged last batch of an epoch)
            eng.freeze_encoder = bool(self.configs.get("linear_eval", False))   # model_utilities.py:160-161: only the head trains
            self.engine = eng
        return eng

    def set_lr(self, lr: float):
        self.lr = float(lr)
        if self.engine is not None:
            self.engine.hp["lr"] = self.lr

    def _to_device(self, batch):
        dev = self.configs["device"]
        b = unpack_batch(batch, self.configs)
        return [stack_inputs(b, self.configs, dev), b["mask"].to(dev, non_blocking=True)]


def train_semantic_segmentation(model, train_loader, val_loader, test_loader, configs, model_configs, process_group=None):
    device = configs["device"]
    model.to(device)
    stepper = FusedSegStepper(model, configs, model_configs, process_group)
    rank0 = _rank(process_group) == 0
    aoi = configs.get("log_AOI_metrics", False)
    metrics = GroupedConfusionMetrics(configs["num_classes"], 3, device, activations=train_loader.dataset.activations if aoi else None)
    sched_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=float(model_configs["learning_rate"]))
    lr_scheduler = init_lr_scheduler(sched_opt, configs, model_configs, steps=len(train_loader))
    best_val, last = 0.0, None
    for epoch in range(configs["epochs"]):
        model.train()
        train_loss = torch.zeros((), dtype=torch.float64, device=device)
        metrics.reset()
        index, loss3 = -1, None
        for index, (batch, nxt) in enumerate(lookahead(train_loader)):
            if nxt is not None:
                stepper.prefetch(nxt)
            loss3, mask = stepper.step_host(batch)
            train_loss += loss3[0].double() * mask.shape[0]
            metrics.update(stepper.engine.pred, mask, activ=batch[-1] if aoi else None)      # segmentation_trainer.py:166-171
        loss_val = float(loss3[0].item()) if index >= 0 else float("nan")
        acc, f1, prec, rec, iou = metrics.compute()
        if configs.get("on_screen_prints"):
            for c in range(3):
                print(f"Train Accuracy ({CLASS_LABELS[c]}): {100 * acc[c].item()}  F-Score: {100 * f1[c].item()}  IoU: {100 * iou[c].item()}")
            print(f"Train MeanIoU: {iou[:3].mean().item() * 100}")
        lr_scheduler.step()
        stepper.set_lr(lr_scheduler.get_last_lr()[0])
        sync_buffers(model, process_group)
        if val_loader is not None:
            val_acc, val_score, miou = eval_semantic_segmentation(model, val_loader, configs, settype="Val", model_configs=model_configs)
            if miou > best_val and configs.get("checkpoint_path") and rank0:
                best_val = miou
                Path(configs["checkpoint_path"]).mkdir(parents=True, exist_ok=True)
                # the reference pickles the whole module (segmentation_trainer.py:255) and main.py:151 torch.load()s it back;
                # __getstate__ of the model mirrors drops the device plans, so the pickle holds parameters and buffers only
                torch.save(model, Path(configs["checkpoint_path"]) / "best_segmentation.pt")
        last = dict(epoch=epoch, loss=loss_val, train_loss=float(train_loss.item()), miou=float(iou[:3].mean().item()))
    return last


def eval_semantic_segmentation(model, loader, configs=None, settype="Val", model_configs=None):
    """segmentation_trainer.py:258-405: eval-mode forward, loss, metrics; returns (100*acc[4], 100*mean F1, 100*mIoU)."""
    device = configs["device"]
    aoi, zones = configs.get("log_AOI_metrics", False), configs.get("log_zone_metrics", False)
    metrics = GroupedConfusionMetrics(configs["num_classes"], 3, device, activations=loader.dataset.activations if aoi else None, zones=zones)
    criterion = create_loss(configs, mode="val")
    model.to(device)
    model.eval()
    total_loss = torch.zeros((), dtype=torch.float64, device=device)
    n = 0
    with torch.no_grad():
        for batch in loader:
            b = unpack_batch(batch, configs)
            image = stack_inputs(b, configs, device)
            mask = b["mask"].to(device, non_blocking=True)
            output = model(image)
            loss = criterion(output, mask)
            pred = getattr(criterion, "last_pred", None)
            predictions = pred if pred is not None else output.argmax(1)
            total_loss += loss.double() * mask.shape[0]
            n += mask.shape[0]
            if predictions.dtype != torch.uint8:
                predictions = predictions.to(torch.uint8)
            metrics.update(predictions.contiguous(), mask, activ=b["activ"] if aoi else None, clz=b["clz"] if zones else None)   # :407-512
    acc, f1, prec, rec, iou = metrics.compute()
    LAST_EVAL[settype] = {"aoi": metrics.compute_aoi(), "zones": metrics.compute_zones(), "samples_per_zone": dict(metrics.samples_per_zone),
                          "confusion": metrics.mat.clone(), "water_fscore": metrics.water_fscore()}
    print(f"{settype} Loss: {(total_loss / max(n, 1)).item()}  MeanIoU: {100 * iou[:3].mean().item()}")
    return 100 * acc, 100 * f1[:3].mean(), 100 * iou[:3].mean()
