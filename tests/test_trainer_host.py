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
