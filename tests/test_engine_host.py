"""Host schedule of the product (kurosiwo_b200/snunet_engine.py) driven through the CPU shadow ops and
checked against the oracle: wiring of the virtual concat, Siamese BN statistics, gradient
first-writer analysis, weight (un)packing.  No GPU needed; the product itself never uses ShadowOps."""
import numpy as np
import pytest
import torch

from kurosiwo_b200.snunet import SNUNet_ECAM
from oracle import snunet_oracle, weights
from shadow_ops import ShadowOps


def _load(model, sd_np):
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})


def test_state_dict_contract():
    m = SNUNet_ECAM(2, 3, base_channel=32)
    sd = weights.make_state(1, 2, 3, 32)
    assert list(m.state_dict().keys()) == list(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    assert sum(p.numel() for p in m.parameters()) == 12034819


def test_cpu_without_backend_fails_loudly():
    m = SNUNet_ECAM(2, 3, base_channel=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 2, 32, 32), torch.zeros(2, 2, 32, 32))


@pytest.mark.parametrize("base,N,H,W,seed", [(8, 2, 32, 32, 11), (8, 3, 48, 32, 5), (32, 2, 32, 32, 12)])
def test_schedule_matches_oracle(base, N, H, W, seed):
    sd_np = weights.make_state(seed, 2, 3, base)
    xA, xB, mask = (torch.from_numpy(a) for a in weights.make_batch(seed, N, H, W))
    sd = snunet_oracle.to_torch_state(sd_np)
    loss_o, logits_o, grads_o = snunet_oracle.train_step(sd, xA, xB, mask)

    model = SNUNet_ECAM(2, 3, base_channel=base, precision="fp32")
    _load(model, sd_np)
    model.set_ops(ShadowOps())
    model.train()
    logits = model(xA, xB)
    loss = snunet_oracle.ce_dice_torch(logits, mask, (1.0, 1.0, 1.0))
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), logits_o.numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(float(loss.detach()), float(loss_o), rtol=1e-4)
    for name, p in model.named_parameters():
        g, go = p.grad, grads_o[name]
        err = (g - go).abs().max().item()
        scale = go.abs().max().item()
        assert err <= 2e-3 * scale + 1e-6, (name, err, scale)
    # running statistics of the shared (Siamese) encoder: updated A then B
    for k in ("conv0_0.bn1.running_mean", "conv0_0.bn1.running_var", "conv4_0.bn2.running_var", "conv0_4.bn2.running_mean"):
        np.testing.assert_allclose(model.state_dict()[k].numpy(), sd[k].numpy(), rtol=1e-4, atol=1e-6)
    assert int(model.state_dict()["conv0_0.bn1.num_batches_tracked"]) == 2
    assert int(model.state_dict()["conv4_0.bn1.num_batches_tracked"]) == 1
    # eval mode uses the running statistics
    model.eval()
    with torch.no_grad():
        ev = model(xA, xB)
        ev_o = snunet_oracle.snunet_forward(sd, xA, xB, training=False)
    np.testing.assert_allclose(ev.numpy(), ev_o.numpy(), rtol=1e-3, atol=2e-4)
