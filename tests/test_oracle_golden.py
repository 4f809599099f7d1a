"""Pins oracle/ against outputs of the UNMODIFIED reference (tests/golden/*.npz, oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle, snunet_oracle, weights


@pytest.mark.parametrize("tag", ["small", "weighted", "ragged", "allignored"])
def test_loss_oracle_matches_reference(golden_dir, tag):
    z = np.load(golden_dir / "loss_cases.npz")
    r = loss_oracle.ce_dice(z[f"{tag}.logits"], z[f"{tag}.labels"], z[f"{tag}.weights"], 3)
    if tag == "allignored":
        assert np.isnan(z[f"{tag}.loss"]) and np.isnan(r["loss"])       # CE of an all-ignored batch is NaN in torch
        np.testing.assert_allclose(r["dice"], z[f"{tag}.dice"], rtol=2e-6)
        return
    np.testing.assert_allclose(r["loss"], z[f"{tag}.loss"], rtol=2e-6)
    np.testing.assert_allclose(r["dice"], z[f"{tag}.dice"], rtol=2e-6)
    np.testing.assert_allclose(r["dlogits"], z[f"{tag}.dlogits"], rtol=2e-4, atol=2e-9)
    np.testing.assert_array_equal(r["argmax"], z[f"{tag}.argmax"])


@pytest.mark.parametrize("tag", ["b8_n2_s32", "b32_n4_s64"])
def test_snunet_oracle_matches_reference(golden_dir, tag):
    fx = np.load(golden_dir / f"snunet_{tag}.npz")
    base, N, H, W, seed = (int(fx[k]) for k in ("base", "N", "H", "W", "seed"))
    sd = snunet_oracle.to_torch_state(weights.make_state(seed, 2, 3, base))
    xA, xB, mask = (torch.from_numpy(a) for a in weights.make_batch(seed, N, H, W))
    loss, logits, grads = snunet_oracle.train_step(sd, xA, xB, mask)
    np.testing.assert_allclose(logits.numpy(), fx["logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(float(loss), float(fx["loss"]), rtol=1e-5)
    names = [str(n) for n in fx["grad_names"]]
    assert names == snunet_oracle.param_names(sd)
    for n, ref_norm in zip(names, fx["grad_norms"]):
        got = float(grads[n].double().norm())
        assert abs(got - ref_norm) <= 1e-3 * ref_norm + 1e-7, (n, got, ref_norm)
    for k in fx.files:
        if k.startswith("grad."):
            g = grads[k[5:]].numpy()
            scale = np.abs(fx[k]).max() + 1e-12
            assert np.abs(g - fx[k]).max() <= 2e-3 * scale + 1e-7, k
        if k.startswith("state."):
            np.testing.assert_allclose(sd[k[6:]].numpy(), fx[k], rtol=1e-4, atol=1e-6)
    with torch.no_grad():
        ev = snunet_oracle.snunet_forward(sd, xA, xB, training=False)
    np.testing.assert_allclose(ev.numpy(), fx["logits_eval"], rtol=1e-4, atol=1e-5)
