"""Host logic of the trainer mirror and of the data-parallel step, on CPU through the shadow ops
(gloo, world_size 2 for the N>1 path).  The product never uses ShadowOps or gloo on its own."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kurosiwo_b200 import change_detection_trainer as cdt
from kurosiwo_b200 import json5lite, synthetic
from kurosiwo_b200.snunet import SNUNet_ECAM
from kurosiwo_b200.utilities import ConfusionMetrics
from oracle import snunet_oracle, weights
from shadow_ops import ShadowOps


def _cfg():
    configs = {"device": "cpu", "inputs": ["pre_event_1", "post_event"], "dem": False, "scale_input": "normalize", "num_classes": 3,
               "loss_function": "ce+dice", "class_weights": [1.0, 1.0, 1.0], "method": "snunet", "epochs": 1, "on_screen_prints": False}
    model_configs = {"method": "snunet", "optimizer": "adam", "learning_rate": 1e-3, "lr_schedule": None, "base_channel": 8}
    return configs, model_configs


def test_json5lite():
    txt = '{\n "a": 1, // c\n "b": [1,2,], /* x */ "s": "http://x//y",\n}'
    assert json5lite.loads(txt) == {"a": 1, "b": [1, 2], "s": "http://x//y"}


def test_confusion_metrics_match_definitions():
    g = torch.Generator().manual_seed(0)
    pred = torch.randint(0, 3, (4, 16, 16), generator=g)
    tgt = torch.randint(0, 4, (4, 16, 16), generator=g)
    m = ConfusionMetrics(3, 3, "cpu")
    m.update(pred[:2], tgt[:2]); m.update(pred[2:].to(torch.uint8), tgt[2:])
    acc, f1, prec, rec, iou = m.compute()
    keep = tgt != 3
    for c in range(3):
        tp = ((pred == c) & (tgt == c) & keep).sum().item()
        fp = ((pred == c) & (tgt != c) & keep).sum().item()
        fn = ((pred != c) & (tgt == c) & keep).sum().item()
        assert abs(acc[c].item() - tp / (tp + fn)) < 1e-12
        assert abs(prec[c].item() - tp / (tp + fp)) < 1e-12
        assert abs(iou[c].item() - tp / (tp + fp + fn)) < 1e-12
        assert abs(f1[c].item() - 2 * tp / (2 * tp + fp + fn)) < 1e-12


def test_train_and_eval_entry_points(monkeypatch, tmp_path):
    configs, model_configs = _cfg()
    configs["checkpoint_path"] = str(tmp_path)
    model = SNUNet_ECAM(2, 3, base_channel=8, precision="fp32")
    model.set_ops(ShadowOps())
    loader = synthetic.SyntheticLoader(2, 2, seed=5, H=32, W=32, pin=False)

    class OracleCrit(torch.nn.Module):
        def forward(self, out, mask):
            return snunet_oracle.ce_dice_torch(out, mask, (1.0, 1.0, 1.0))
    monkeypatch.setattr(cdt, "create_loss", lambda configs, mode="val": OracleCrit())
    last = cdt.train_change_detection(model, loader, loader, loader, configs, model_configs)
    assert np.isfinite(last["loss"]) and 0.0 <= last["miou"] <= 1.0
    assert (tmp_path / "checkpoint_epoch=0.pt").exists()
    ck = torch.load(tmp_path / "checkpoint_epoch=0.pt", weights_only=False)
    assert set(ck) >= {"epoch", "model_state_dict", "optimizer_state_dict", "lr_scheduler_state_dict", "loss"}
    acc, f1, miou = cdt.eval_change_detection(model, loader, "Validation", configs, model_configs)
    assert acc.shape == (4,) and 0 <= float(miou) <= 100
    # first step of the fused path equals the oracle's first step
    sd = snunet_oracle.to_torch_state({k: v.detach().numpy() for k, v in SNUNet_ECAM(2, 3, 8).state_dict().items()})


def test_fused_step_matches_oracle_with_adam():
    seed, base, N, H, W = 3, 8, 2, 32, 32
    sd_np = weights.make_state(seed, 2, 3, base)
    xA, xB, mask = (torch.from_numpy(a) for a in weights.make_batch(seed, N, H, W))
    sd = snunet_oracle.to_torch_state(sd_np)
    model = SNUNet_ECAM(2, 3, base_channel=base, precision="fp32")
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
    model.set_ops(ShadowOps())
    eng = model.engine(xA)
    eng.init_training(lr=1e-3)
    state = {}
    for it in range(2):
        loss_o, _, grads = snunet_oracle.train_step(sd, xA, xB, mask)
        snunet_oracle.adam_step(sd, grads, state, lr=1e-3)
        l3 = eng.train_step(xA, xB, mask)
        np.testing.assert_allclose(l3[0].item(), float(loss_o), rtol=1e-3)
        assert torch.equal(eng.pred.long(), eng.logits.argmax(1))
    for n, p in model.named_parameters():
        assert (p.detach() - sd[n]).abs().max().item() < 2e-3, n


def _ddp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    seed, base, N, H, W = 7, 8, 2, 32, 32
    sd_np = weights.make_state(seed, 2, 3, base)
    if rank == 1:   # rank 1 starts from different weights: the initial broadcast must overwrite them
        sd_np = {k: (v + 0.01 if v.dtype == np.float32 and not k.endswith(("running_mean", "running_var")) else v) for k, v in sd_np.items()}
    model = SNUNet_ECAM(2, 3, base_channel=base, precision="fp32")
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
    model.set_ops(ShadowOps())
    xA, xB, mask = (torch.from_numpy(a) for a in weights.make_batch(seed + rank, N, H, W))
    eng = model.engine(xA)
    eng.init_training(lr=1e-3, process_group=dist.group.WORLD)
    eng.train_step(xA, xB, mask)
    ret[rank] = (eng.params.flat.clone(), eng.params.grad.clone())
    dist.destroy_process_group()


def test_data_parallel_step_gloo_world2():
    world, port = 2, 29611
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, port, ret), nprocs=world, join=True)
    p0, g0 = ret[0]
    p1, g1 = ret[1]
    assert torch.equal(p0, p1) and torch.equal(g0, g1)          # identical replicas after the step
    # the all-reduced gradient is the SUM of the two ranks' gradients (scaled by 1/world inside Adam)
    seed, base, N, H, W = 7, 8, 2, 32, 32
    tot = None
    for r in range(2):
        sd = snunet_oracle.to_torch_state(weights.make_state(seed, 2, 3, base))
        xA, xB, mask = (torch.from_numpy(a) for a in weights.make_batch(seed + r, N, H, W))
        _, _, grads = snunet_oracle.train_step(sd, xA, xB, mask)
        tot = grads if tot is None else {k: tot[k] + grads[k] for k in grads}
    model = SNUNet_ECAM(2, 3, base_channel=base, precision="fp32")
    from kurosiwo_b200.snunet_engine import FlatParams
    fp = FlatParams(model)
    for n in ("conv0_0.conv1.weight", "conv2_1.conv1.weight", "conv_final.bias"):
        off, shape = fp.offsets[n]
        got = g0[off:off + shape.numel()].view(shape)
        assert (got - tot[n]).abs().max().item() <= 2e-3 * tot[n].abs().max().item() + 1e-6, n


def test_optimizer_state_survives_a_ragged_batch():
    """A batch of another size builds a new engine (new activation buffers); the Adam moments and step count must carry over."""
    configs, model_configs = _cfg()
    model = SNUNet_ECAM(2, 3, base_channel=8, precision="fp32")
    model.set_ops(ShadowOps())
    stepper = cdt.FusedStepper(model, configs, model_configs)
    full = synthetic.make_batch(11, 2, 32, 32, False)
    ragged = synthetic.make_batch(12, 1, 32, 32, False)
    stepper.step_host(full); stepper.step_host(full)
    e1 = stepper.engine
    m_before = e1.adam_m.clone()
    assert int(e1.adam_step.item()) == 2 and float(m_before.abs().sum()) > 0
    stepper.step_host(ragged)
    e2 = stepper.engine
    assert e2 is not e1 and int(e2.adam_step.item()) == 3
    # after one more step the first moment is 0.9 * old + 0.1 * g: it cannot be the fresh-state value 0.1 * g
    g = e2.params.grad
    assert not torch.allclose(e2.adam_m, 0.1 * g, rtol=1e-3, atol=1e-9)
    assert torch.allclose(e2.adam_m, 0.9 * m_before + 0.1 * g, rtol=1e-4, atol=1e-8)


def test_lookahead_pairs():
    from kurosiwo_b200.host_pipeline import lookahead
    assert list(lookahead([])) == []
    assert list(lookahead([1])) == [(1, None)]
    assert list(lookahead("abc")) == [("a", "b"), ("b", "c"), ("c", None)]


# ---- round 2: grouped metrics, AdamW, state hand-over, best-checkpoint reload, reference multi_scale_train ---------------------
def test_grouped_confusion_metrics_match_per_group_definitions():
    """Per-activation (AOI) and per-climate-zone metric sets (reference change_detection_trainer.py:184-199, :445-480) and the
    water-only F-score (:408-413): the grouped matrices must equal the global definition applied to each group's samples."""
    from kurosiwo_b200.utilities import GroupedConfusionMetrics, metrics_from_confusion, water_only_fscore
    g = torch.Generator().manual_seed(1)
    B = 6
    pred = torch.randint(0, 3, (B, 12, 10), generator=g).to(torch.uint8)
    tgt = torch.randint(0, 4, (B, 12, 10), generator=g)
    acts = [130, 470, 555, 1111011]
    activ = torch.tensor([470, 130, 470, 1111011, 999, 130])        # 999: not a known activation -> only the global set
    clz = torch.tensor([1, 2, 2, 3, 1, 2])
    m = GroupedConfusionMetrics(3, 3, "cpu", activations=acts, zones=True)
    m.update(pred[:4], tgt[:4], activ=activ[:4], clz=clz[:4])
    m.update(pred[4:], tgt[4:], activ=activ[4:], clz=clz[4:])
    ref = ConfusionMetrics(3, 3, "cpu"); ref.update(pred, tgt)
    assert torch.equal(m.mat, ref.mat)
    aoi = m.compute_aoi()
    assert sorted(aoi) == [130, 470, 1111011]
    for a in aoi:
        r = ConfusionMetrics(3, 3, "cpu"); r.update(pred[activ == a], tgt[activ == a])
        for got, want in zip(aoi[a], r.compute()):
            assert torch.allclose(got, want)
    zones = m.compute_zones()
    assert m.samples_per_zone == {1: 2, 2: 3, 3: 1}
    for z in (1, 2, 3):
        r = ConfusionMetrics(3, 3, "cpu"); r.update(pred[clz == z], tgt[clz == z])
        assert torch.allclose(zones[z][4], r.compute()[4])
    # water-only F1 = F1 of the relabelled 2-class problem (2 -> 1 in predictions and labels, ignore_index 3)
    p2, t2 = pred.long().clone(), tgt.clone()
    p2[p2 == 2] = 1; t2[t2 == 2] = 1
    keep = t2 != 3
    tp = ((p2 == 1) & (t2 == 1) & keep).sum().item(); fp = ((p2 == 1) & (t2 == 0) & keep).sum().item(); fn = ((p2 == 0) & (t2 == 1) & keep).sum().item()
    assert abs(water_only_fscore(m.mat)[1].item() - 2 * tp / (2 * tp + fp + fn)) < 1e-12
    assert torch.allclose(metrics_from_confusion(m.mat)[4], ref.compute()[4])


def test_adamw_config_reaches_the_fused_optimizer():
    """optimizer 'adamw' (reference change_detection_trainer.py:55-60: betas + weight_decay from the method config) against
    torch.optim.AdamW driven with the oracle's gradients."""
    seed, base, N, H, W = 3, 8, 2, 32, 32
    configs, model_configs = _cfg()
    model_configs.update(optimizer="adamw", betas=[0.8, 0.95], weight_decay=0.05, learning_rate=2e-3)
    sd_np = weights.make_state(seed, 2, 3, base)
    model = SNUNet_ECAM(2, 3, base_channel=base, precision="fp32")
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
    model.set_ops(ShadowOps())
    stepper = cdt.FusedStepper(model, configs, model_configs)
    batch = synthetic.make_batch(21, N, H, W, False)
    xA, xB, mask = stepper._to_device(batch)
    sd = snunet_oracle.to_torch_state(sd_np)
    names = [n for n, _ in model.named_parameters()]
    ref_params = [torch.nn.Parameter(sd[n].clone()) for n in names]
    opt = torch.optim.AdamW(ref_params, lr=2e-3, betas=(0.8, 0.95), weight_decay=0.05)
    for _ in range(2):
        _, _, grads = snunet_oracle.train_step({**sd, **{n: p.detach() for n, p in zip(names, ref_params)}}, xA, xB, mask)
        for n, p in zip(names, ref_params):
            p.grad = grads[n].clone()
        opt.step()
        stepper.step_host(batch)
    assert stepper.engine.optimizer == "adamw" and stepper.engine.hp["wd"] == 0.05 and stepper.engine.hp["b1"] == 0.8
    for n, p in zip(names, ref_params):
        got = dict(model.named_parameters())[n].detach()
        assert (got - p.detach()).abs().max().item() < 3e-3, n     # Adam-type updates are lr-sized: 2 steps x 2e-3, rounding-level gradient noise


def test_engine_swap_hands_over_moments_step_and_dropout_counter():
    """assert the hand-over itself (not two runs that would lose the state identically)."""
    from kurosiwo_b200.siam_unet import SiamUnet_conc
    configs, model_configs = _cfg()
    configs["method"] = "siam-conc"
    model = SiamUnet_conc(input_nbr=2, label_nbr=3, precision="fp32")
    model.set_ops(ShadowOps())
    stepper = cdt.FusedStepper(model, configs, model_configs)
    full, ragged = synthetic.make_batch(11, 2, 32, 32, False), synthetic.make_batch(12, 1, 32, 32, False)
    stepper.step_host(full); stepper.step_host(full)
    e1 = stepper.engine
    m1, v1, s1, d1 = e1.adam_m.clone(), e1.adam_v.clone(), int(e1.adam_step.item()), int(e1.do_step.item())
    assert s1 == 2 and d1 == 2 and float(m1.abs().sum()) > 0
    xa = stepper._to_device(ragged)[0]
    e2 = stepper._engine(xa)                       # the swap, before any step on the new engine
    assert e2 is not e1
    assert torch.equal(e2.adam_m, m1) and torch.equal(e2.adam_v, v1) and int(e2.adam_step.item()) == s1
    assert int(e2.do_step.item()) == d1            # the dropout mask sequence continues instead of restarting


def test_reference_multi_scale_train_raises():
    """The reference's multi_scale_train branch (change_detection_trainer.py:155-164) resizes the [B,H,W] int64 mask with
    F.interpolate(mask, size=int, mode="nearest"): torch has no nearest kernel for Long, so the branch raises before any loss is
    computed (and a float mask would be resized along W only: [B,H,W] is read as (N, C, L)).  The drop-in raises the same type."""
    import torch.nn.functional as F
    mask = torch.randint(0, 4, (2, 64, 64))
    with pytest.raises(NotImplementedError):
        F.interpolate(mask, size=16, mode="nearest")
    assert F.interpolate(mask.float(), size=16, mode="nearest").shape == (2, 64, 16)
    from kurosiwo_b200.changeformer import ChangeFormerV6
    configs, model_configs = _cfg()
    model_configs["multi_scale_train"] = True
    with pytest.raises(NotImplementedError):
        cdt.FusedStepper(ChangeFormerV6(input_nc=2, output_nc=3, decoder_softmax=True, embed_dim=32), configs, model_configs)


def test_main_reloads_best_checkpoint(tmp_path):
    """the test set is evaluated on best_segmentation.pt (reference main.py:149-152, :172-185), not on the last epoch."""
    import importlib.util
    import sys
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("ks_main", Path(__file__).resolve().parent.parent / "main.py")
    main = importlib.util.module_from_spec(spec); spec.loader.exec_module(main)
    configs = {"checkpoint_path": str(tmp_path), "device": "cpu"}
    best = SNUNet_ECAM(2, 3, base_channel=8, precision="fp32")
    torch.save({"epoch": 0, "model_state_dict": best.state_dict(), "optimizer_state_dict": {}, "lr_scheduler_state_dict": {}, "loss": 0.0},
               tmp_path / "best_segmentation.pt")
    last = SNUNet_ECAM(2, 3, base_channel=8, precision="fp32")
    assert not torch.equal(last.conv0_0.conv1.weight, best.conv0_0.conv1.weight)
    out = main.load_best_checkpoint(last, configs, "cd")
    assert out is last and torch.equal(last.conv0_0.conv1.weight, best.conv0_0.conv1.weight)
    (tmp_path / "best_segmentation.pt").unlink()
    torch.save(best, tmp_path / "best_segmentation.pt")               # the segmentation trainer pickles the module
    out = main.load_best_checkpoint(last, configs, "segmentation")
    assert isinstance(out, SNUNet_ECAM) and out is not last
    (tmp_path / "best_segmentation.pt").unlink()
    assert main.load_best_checkpoint(last, configs, "cd") is last      # no file: current weights


def test_best_checkpoint_has_the_reference_keys(tmp_path):
    configs, model_configs = _cfg()
    configs["checkpoint_path"] = str(tmp_path)
    configs["log_AOI_metrics"] = True
    configs["log_zone_metrics"] = True
    configs["evaluate_water"] = True
    model = SNUNet_ECAM(2, 3, base_channel=8, precision="fp32")
    model.set_ops(ShadowOps())
    loader = synthetic.SyntheticLoader(2, 2, seed=5, H=32, W=32, pin=False)

    class OracleCrit(torch.nn.Module):
        def forward(self, out, mask):
            return snunet_oracle.ce_dice_torch(out, mask, (1.0, 1.0, 1.0))
    import kurosiwo_b200.change_detection_trainer as mod
    orig = mod.create_loss
    mod.create_loss = lambda c, mode="val": OracleCrit()
    try:
        cdt.train_change_detection(model, loader, loader, loader, configs, model_configs)
    finally:
        mod.create_loss = orig
    ck = torch.load(tmp_path / "best_segmentation.pt", weights_only=False)
    assert set(ck) >= {"epoch", "model_state_dict", "optimizer_state_dict", "lr_scheduler_state_dict", "loss"}   # reference :312-318
    det = cdt.LAST_EVAL["Validation"]
    assert det["zones"] and det["aoi"] and "water_fscore" in det
    assert sum(det["samples_per_zone"].values()) == 4


def _ddp_vit_worker(rank, world, port, ret, bucket_mb):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from kurosiwo_b200.vision_transformer import FinetunerSegmentation, ViT
    from oracle import vit_oracle
    dim, depth, heads, mlp = 64, 3, 2, 128
    sd_np = vit_oracle.make_state(5, dim, depth, heads, mlp)
    enc = ViT(image_size=224, patch_size=16, num_classes=3, dim=dim, depth=depth, heads=heads, mlp_dim=mlp, channels=6, precision="fp32")
    model = FinetunerSegmentation(encoder=enc, configs={"mlp": False, "decoder": False, "num_classes": 3, "finetuning_patch_size": 16})
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
    model.set_ops(ShadowOps())
    img, mask = (torch.from_numpy(a) for a in vit_oracle.make_batch(5 + rank, 1))
    eng = model.engine(img)
    eng.init_training(lr=1e-3, process_group=dist.group.WORLD, bucket_mb=bucket_mb)
    eng._fwd_loss_bwd(img, mask)
    local = eng.params.grad.clone()              # NOT yet complete: buckets announced during the backward are already reduced
    eng._allreduce()
    ret[rank] = (eng.params.grad.clone(), dict(eng.comm_stats), local)
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_mb", [0.05, 1e9])
def test_bucketed_allreduce_gloo_world2(bucket_mb):
    """The bucketed exchange (buckets announced block by block during the ViT backward; the rest after it) must give every rank the
    SUM of the ranks' gradients, whatever the bucket size - against the oracle's per-rank gradients."""
    from oracle import vit_oracle
    world, port = 2, 29633 + (0 if bucket_mb < 1 else 1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_vit_worker, args=(world, port, ret, bucket_mb), nprocs=world, join=True)
    g0, st0, _ = ret[0]
    g1, st1, _ = ret[1]
    assert torch.equal(g0, g1)
    total_bytes = 4 * g0.numel()
    assert st0["bytes"] == total_bytes                           # every element exactly once
    assert (st0["messages"] > 2) == (bucket_mb < 1)              # small buckets: several messages; huge bucket: one (plus none)
    dim, depth, heads, mlp = 64, 3, 2, 128
    tot = None
    for r in range(2):
        sd = vit_oracle.to_torch_state(vit_oracle.make_state(5, dim, depth, heads, mlp))
        img, mask = (torch.from_numpy(a) for a in vit_oracle.make_batch(5 + r, 1))
        _, _, grads = vit_oracle.train_step(sd, img, mask, heads)
        tot = grads if tot is None else {k: tot[k] + grads[k] for k in grads}
    from kurosiwo_b200.engine_common import FlatParams
    from kurosiwo_b200.vision_transformer import FinetunerSegmentation, ViT
    enc = ViT(image_size=224, patch_size=16, num_classes=3, dim=dim, depth=depth, heads=heads, mlp_dim=mlp, channels=6, precision="fp32")
    fp = FlatParams(FinetunerSegmentation(encoder=enc, configs={"mlp": False, "decoder": False, "num_classes": 3, "finetuning_patch_size": 16}))
    for n in ("model.transformer.layers.0.0.to_qkv.weight", "model.transformer.layers.2.1.net.4.weight", "head.weight", "model.pos_embedding"):
        off, shape = fp.offsets[n]
        got = g0[off:off + shape.numel()].view(shape)
        assert (got - tot[n]).abs().max().item() <= 2e-3 * tot[n].abs().max().item() + 1e-6, n
