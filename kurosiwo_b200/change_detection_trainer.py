"""Mirror of the reference's `training/change_detection_trainer.py` entry points for the B200 path.

`train_change_detection(model, train_loader, val_loader, test_loader, configs, model_configs)` and
`eval_change_detection(model, loader, settype, configs, model_configs)` keep the reference signatures,
batch tuple layout (dataset/Dataset.py:826-839), checkpoint dict layout (:206-213) and return values (:791).
The inner loop (:93-204) is replaced: pinned H2D copies, ONE fused engine step (forward, CE+Dice+argmax,
backward, Adam) with no per-iteration host sync, metrics from one device-side confusion matrix.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch

from .changeformer import ChangeFormerV6
from .siam_unet import _SiamUnet
from .snunet import SNUNet_ECAM
from .host_pipeline import HostPipelineMixin, lookahead
from .utilities import ConfusionMetrics, GroupedConfusionMetrics, create_loss, init_lr_scheduler

CLASS_LABELS = {0: "No water", 1: "Permanent Waters", 2: "Floods", 3: "Invalid pixels"}


def unpack_batch(batch, configs):
    """change_detection_trainer.py:95-106 -> dict(post_event, mask, pre_event_1, pre_event_2, dem, clz, activ)."""
    if configs.get("scale_input") is not None:
        if configs.get("dem"):
            (_, _, post, mask, _, _, pre1, _, _, pre2, dem, clz, activ) = batch
        else:
            (_, _, post, mask, _, _, pre1, _, _, pre2, clz, activ) = batch
            dem = None
    else:
        if configs.get("dem"):
            post, mask, pre1, pre2, dem, clz, activ = batch
        else:
            post, mask, pre1, pre2, clz, activ = batch
            dem = None
    return dict(post_event=post, mask=mask, pre_event_1=pre1, pre_event_2=pre2, dem=dem, clz=clz, activ=activ)


_PRE = {}


def preprocess_raw(x, configs):
    """`raw_input: true` (additive key): the loader yields RAW float32 SAR tiles (what cv.imread returns, dataset/Dataset.py:720-760)
    and the reference's per-sample CPU transform - clamp to [0, clamp_input], nan_to_num, Normalize(data_mean, data_std)
    (dataset/Dataset.py:162-168, :192-198) - runs as ONE kernel on the device copy, in place."""
    if not configs.get("raw_input") or not x.is_cuda:
        return x
    from .lib import default_ops
    key = (str(x.device), tuple(configs.get("data_mean", ())), tuple(configs.get("data_std", ())))
    if key not in _PRE:
        _PRE[key] = (torch.tensor(configs["data_mean"], dtype=torch.float32, device=x.device),
                     torch.tensor(configs["data_std"], dtype=torch.float32, device=x.device))
    mean, std = _PRE[key]
    x = x.contiguous()
    default_ops().sar_preprocess(x, x, mean, std, configs.get("clamp_input") or 0.0)
    return x


def select_inputs(b, configs, device):
    """change_detection_trainer.py:117-133: the two images in `configs['inputs']` order (+DEM channel)."""
    outs = []
    for name in configs["inputs"]:
        x = preprocess_raw(b[name].to(device, non_blocking=True), configs)
        if configs.get("dem"):
            x = torch.cat((x, b["dem"].to(device, non_blocking=True)), dim=1)
        outs.append(x)
    return outs


class FusedStepper(HostPipelineMixin):
    """Owns the engine-side training state for one model/batch geometry (the public fast path)."""

    def __init__(self, model, configs, model_configs, process_group=None):
        if not isinstance(model, (SNUNet_ECAM, _SiamUnet, ChangeFormerV6)):
            raise TypeError("the fused step is implemented for kurosiwo_b200's SNUNet_ECAM, SiamUnet_conc/diff and ChangeFormerV6")
        if configs.get("loss_function", "ce+dice") not in ("ce+dice", "cross_entropy"):
            raise NotImplementedError("the fused step computes CE+Dice (utilities/bce_and_dice.py) or plain cross-entropy "
                                      "(utilities/utilities.py:308-321); other losses are outside the B200 hot path")
        opt = model_configs.get("optimizer", "adam")
        if opt not in ("adam", "adamw", "sgd"):
            raise ValueError(f"optimizer '{opt}': the reference knows 'adam' (change_detection_trainer.py:52-54), 'adamw' (:55-60), 'sgd' (:61-66)")
        if model_configs.get("multi_scale_train"):
            # The reference's branch (:155-164) calls F.interpolate(mask, size=int, mode="nearest") on the 3-D int64 mask: torch raises
            # NotImplementedError('"compute_indices_weights_nearest" not implemented for \'Long\'') - pinned by
            # tests/test_trainer_host.py::test_reference_multi_scale_train_raises.  A drop-in keeps that behaviour.
            raise NotImplementedError("multi_scale_train: the reference branch (change_detection_trainer.py:155-164) raises "
                                      "NotImplementedError itself (F.interpolate of the int64 mask); it is kept unusable here as well")
        self.opt = opt
        self.model, self.configs, self.model_configs, self.pg = model, configs, model_configs, process_group
        self.engine = None
        self.lr = float(model_configs["learning_rate"])

    def _engine(self, x):
        eng = self.model.engine(x)
        if eng is not self.engine:
            eng.init_training(class_weights=self.configs.get("class_weights", [1.0, 1.0, 1.0]), ignore_index=3, lr=self.lr,
                              # the reference's Adam gets only lr (:52-54); AdamW gets betas + weight_decay (:55-60); SGD momentum + weight_decay (:61-66)
                              betas=tuple(self.model_configs["betas"]) if self.opt == "adamw" else (0.9, 0.999), eps=1e-8,
                              weight_decay=float(self.model_configs.get("weight_decay", 0.0)) if self.opt in ("sgd", "adamw") else 0.0,
                              optimizer=self.opt, momentum=float(self.model_configs.get("momentum", 0.0)),
                              process_group=self.pg,
                              dice_weight=0.0 if self.configs.get("loss_function") == "cross_entropy" else 1.0)
            eng.adopt_training_state(self.engine)   # another batch geometry (e.g. the ragged last batch of an epoch)
            self.engine = eng
        return eng

    def set_lr(self, lr: float):
        self.lr = float(lr)
        if self.engine is not None:
            self.engine.hp["lr"] = self.lr

    def _to_device(self, batch):
        dev = self.configs["device"]
        b = unpack_batch(batch, self.configs)
        xa, xb = select_inputs(b, self.configs, dev)
        return [xa, xb, b["mask"].to(dev, non_blocking=True)]


def _train_predictions(engine, model_configs):
    """The class map the training metrics see.  Default: the argmax the loss kernel emitted.  ChangeFormer with
    `multi_scale_infer` (change_detection_trainer.py:139-148): argmax of the mean of the five outputs, the coarser ones resized with
    mode='nearest' - a metrics-only side computation on the outputs the engine already holds (off by default in the reference)."""
    if model_configs.get("multi_scale_infer") and hasattr(engine, "outputs"):
        outs = engine.outputs()
        size = outs[-1].shape[2]
        final = torch.zeros_like(outs[-1], dtype=torch.float32)
        for o in outs:
            o = o.float()
            final += torch.nn.functional.interpolate(o, size=size, mode="nearest") if o.shape[2] != size else o
        return (final / len(outs)).argmax(1).to(torch.uint8)
    return engine.pred


def _rank(process_group) -> int:
    if process_group is None:
        return 0
    import torch.distributed as dist
    return dist.get_rank(process_group)


def sync_buffers(model, process_group):
    """Data parallelism keeps BatchNorm running statistics rank-local during training (torch DDP without SyncBN broadcasts rank 0's
    buffers at every forward; here ONCE, right before they are read: evaluation and checkpoints) - so that every rank evaluates and
    saves the same model."""
    if process_group is None:
        return
    import torch.distributed as dist
    src = dist.get_global_rank(process_group, 0) if hasattr(dist, "get_global_rank") else 0
    for b in model.buffers():
        if b.numel():
            dist.broadcast(b, src=src, group=process_group)


def _optimizer_state(stepper):
    eng = stepper.engine
    return {"kind": stepper.opt if hasattr(stepper, "opt") else "adam", "exp_avg": eng.adam_m, "exp_avg_sq": eng.adam_v, "step": eng.adam_step,
            "hyper": dict(eng.hp), "layout": {n: (o, tuple(s)) for n, (o, s) in eng.params.offsets.items()}}


def train_change_detection(model, train_loader, val_loader, test_loader, configs, model_configs, process_group=None):
    assert len(configs["inputs"]) == 2, f'Model {model_configs.get("method")} requires exactly 2 input images.'
    device = configs["device"]
    model.to(device)
    stepper = FusedStepper(model, configs, model_configs, process_group)
    rank0 = _rank(process_group) == 0
    aoi = configs.get("log_AOI_metrics", False)
    metrics = GroupedConfusionMetrics(configs["num_classes"], 3, device, activations=train_loader.dataset.activations if aoi else None)
    # a torch optimizer object only to drive the reference's per-epoch LR scheduler / checkpoint layout
    sched_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=float(model_configs["learning_rate"]))
    lr_scheduler = init_lr_scheduler(sched_opt, configs, model_configs, steps=len(train_loader))
    best_val, last = 0.0, None
    print(f'===== checkpoint_path: {configs.get("checkpoint_path")} ====')
    for epoch in range(0, configs["epochs"]):
        model.train()
        train_loss = torch.zeros((), dtype=torch.float64, device=device)
        metrics.reset()
        index = -1
        for index, (batch, nxt) in enumerate(lookahead(train_loader)):
            if nxt is not None:
                stepper.prefetch(nxt)                                # H2D of the next batch overlaps this step
            loss3, mask = stepper.step_host(batch)
            train_loss += loss3[0].double() * mask.shape[0]          # stays on device: no .item() in the loop
            metrics.update(_train_predictions(stepper.engine, model_configs), mask, activ=batch[-1] if aoi else None)   # :184-199
            if configs.get("on_screen_prints") and index % configs.get("print_frequency", 10) == 0:
                print(f"({epoch}) it {index} Train Loss: {train_loss.item():.4f}")
        loss_val = float(loss3[0].item()) if index >= 0 else float("nan")
        sync_buffers(model, process_group)
        if rank0 and configs.get("checkpoint_path") and index % configs.get("train_save_checkpoint_freq", 1) == 0:
            torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": _optimizer_state(stepper),
                        "lr_scheduler_state_dict": lr_scheduler.state_dict(), "loss": loss_val},      # the reference's keys (:206-213)
                       Path(configs["checkpoint_path"]) / f"checkpoint_epoch={epoch}.pt")
        acc, f1, prec, rec, iou = metrics.compute()
        if configs.get("on_screen_prints"):
            for c in range(3):
                print(f"Train Accuracy ({CLASS_LABELS[c]}): {100 * acc[c].item()}  F-Score: {100 * f1[c].item()}  IoU: {100 * iou[c].item()}")
            print(f"Train MeanIoU: {iou[:3].mean().item() * 100}")
            for a, (a_acc, a_f1, _, _, a_iou) in metrics.compute_aoi().items():
                print(f"Train AOI {a}: accuracy {[round(100 * x, 3) for x in a_acc[:3].tolist()]} IoU {[round(100 * x, 3) for x in a_iou[:3].tolist()]}")
        lr_scheduler.step()
        stepper.set_lr(lr_scheduler.get_last_lr()[0])
        if val_loader is not None:
            val_acc, val_score, miou = eval_change_detection(model, val_loader, settype="Validation", configs=configs, model_configs=model_configs)
            if miou > best_val and configs.get("checkpoint_path") and rank0:
                best_val = miou
                torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": _optimizer_state(stepper),
                            "lr_scheduler_state_dict": lr_scheduler.state_dict(), "loss": loss_val},  # the reference's keys (:312-318)
                           Path(configs["checkpoint_path"]) / "best_segmentation.pt")
                (Path(configs["checkpoint_path"]) / "best_segmentation.txt").write_text(f"{epoch}\n{miou}")
        last = dict(epoch=epoch, loss=loss_val, train_loss=float(train_loss.item()), miou=float(iou[:3].mean().item()))
    return last


LAST_EVAL = {}     # details of the most recent eval_*: per-AOI / per-zone metric sets, water-only F-score (the reference logs them to wandb)


def eval_change_detection(model, loader, settype, configs=None, model_configs=None):
    """change_detection_trainer.py:325-791: eval-mode forward (running-stat BN), CE(+Dice) loss, metrics: global, per climate zone
    (`log_zone_metrics`, :445-470), per activation (`log_AOI_metrics`, :472-480) and water-only F-score (`evaluate_water`, :408-413).
    Returns (100*accuracy[4], 100*mean F1, 100*mIoU) like the reference (:791); the grouped sets land in LAST_EVAL[settype]."""
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
            inputs = select_inputs(b, configs, device)
            mask = b["mask"].to(device, non_blocking=True)
            output = model(*inputs)
            if isinstance(output, (list, tuple)):       # ChangeFormer returns [p_c4, p_c3, p_c2, p_c1, cp]; the last is used (:396)
                output = output[-1]
            loss = criterion(output, mask)
            pred = getattr(criterion, "last_pred", None)
            predictions = pred if pred is not None else output.argmax(1)
            total_loss += loss.double() * mask.shape[0]
            n += mask.shape[0]
            if predictions.dtype != torch.uint8:
                predictions = predictions.to(torch.uint8)
            metrics.update(predictions.contiguous(), mask, activ=b["activ"] if aoi else None, clz=b["clz"] if zones else None)
    acc, f1, prec, rec, iou = metrics.compute()
    details = {"aoi": metrics.compute_aoi(), "zones": metrics.compute_zones(), "samples_per_zone": dict(metrics.samples_per_zone),
               "confusion": metrics.mat.clone()}
    print(f"{settype} Loss: {(total_loss / max(n, 1)).item()}  MeanIoU: {100 * iou[:3].mean().item()}")
    if configs.get("evaluate_water"):
        details["water_fscore"] = metrics.water_fscore()
        print(f"{settype} F-Score (Only water): {100 * details['water_fscore'][1].item()}")
    for z, (z_acc, z_f1, _, _, z_iou) in details["zones"].items():
        print(f"{settype} climate zone {z}: MeanIoU {100 * z_iou[:3].mean().item()} ({details['samples_per_zone'][z]} samples)")
    LAST_EVAL[settype] = details
    return 100 * acc, 100 * f1[:3].mean(), 100 * iou[:3].mean()
