"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python -m oracle.make_golden
The fixtures pin oracle/ (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    sys.path.insert(0, REF)
    from models.snunet import SNUNet_ECAM as RefSNUNet          # noqa: E402  (reference, read-only)
    from utilities.bce_and_dice import BCEandDiceLoss as RefLoss  # noqa: E402
    from oracle.weights import make_batch, make_state

    OUT.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---- loss fixtures ------------------------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(7))
    cases = {}
    for tag, (N, H, W, wts, all_ignored) in {
        "small": (2, 16, 16, [1.0, 1.0, 1.0], False),
        "weighted": (3, 12, 20, [0.3715753140309927, 14.009780283125977, 8.20405370357821], False),
        "ragged": (1, 7, 9, [1.0, 2.0, 0.5], False),
        "allignored": (2, 8, 8, [1.0, 1.0, 1.0], True),
    }.items():
        z = (3.0 * rng.standard_normal((N, 3, H, W))).astype(np.float32)
        y = rng.choice(4, size=(N, H, W), p=[0.6, 0.1, 0.15, 0.15]).astype(np.int64)
        if all_ignored:
            y[:] = 3
        zt = torch.from_numpy(z).requires_grad_(True)
        crit = RefLoss(weights=torch.tensor(wts), ignore_index=3, use_softmax=True)
        loss = crit(zt, torch.from_numpy(y))
        dice = crit.dice(zt, torch.from_numpy(y))
        loss.backward()
        cases[tag] = dict(logits=z, labels=y, weights=np.array(wts, np.float32), loss=loss.detach().numpy(),
                          dice=dice.detach().numpy(), dlogits=zt.grad.numpy(), argmax=zt.detach().argmax(1).numpy().astype(np.uint8))
    np.savez_compressed(OUT / "loss_cases.npz", **{f"{t}.{k}": v for t, c in cases.items() for k, v in c.items()})

    # ---- SNUNet fixtures: forward logits, loss, gradients, running stats after one step ----
    for tag, (base, N, H, W, seed) in {"b8_n2_s32": (8, 2, 32, 32, 11), "b32_n4_s64": (32, 4, 64, 64, 12)}.items():
        sd = make_state(seed, 2, 3, base)
        xA, xB, mask = make_batch(seed, N, H, W)
        model = RefSNUNet(2, 3, base_channel=base)
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
        model.train()
        crit = RefLoss(weights=torch.tensor([1.0, 1.0, 1.0]), ignore_index=3, use_softmax=True)
        out = model(torch.from_numpy(xA), torch.from_numpy(xB))
        loss = crit(out, torch.from_numpy(mask))
        loss.backward()
        fx = {"base": base, "N": N, "H": H, "W": W, "seed": seed, "logits": out.detach().numpy(), "loss": loss.detach().numpy()}
        keep_full = {"conv0_0.conv1.weight", "conv0_0.conv1.bias", "conv0_0.bn1.weight", "conv0_0.bn2.bias", "conv0_4.conv1.weight",
                     "conv1_2.conv2.weight", "Up1_0.up.weight", "Up2_1.up.bias", "ca.fc1.weight", "ca.fc2.weight", "ca1.fc1.weight",
                     "ca1.fc2.weight", "conv_final.weight", "conv_final.bias", "conv3_0.bn1.weight", "conv2_1.conv1.bias"}
        names, norms = [], []
        for k, p in model.named_parameters():
            names.append(k)
            norms.append(float(p.grad.double().norm()))
            if k in keep_full:
                fx[f"grad.{k}"] = p.grad.numpy()
        fx["grad_names"] = np.array(names)
        fx["grad_norms"] = np.array(norms, np.float64)
        st = model.state_dict()
        for k in ("conv0_0.bn1.running_mean", "conv0_0.bn1.running_var", "conv0_0.bn1.num_batches_tracked",
                  "conv4_0.bn2.running_var", "conv0_4.bn2.running_mean", "conv0_4.bn1.num_batches_tracked"):
            fx[f"state.{k}"] = st[k].numpy()
        # eval-mode forward with the updated running stats
        model.eval()
        with torch.no_grad():
            fx["logits_eval"] = model(torch.from_numpy(xA), torch.from_numpy(xB)).numpy()
        np.savez_compressed(OUT / f"snunet_{tag}.npz", **fx)
        print(tag, "loss", float(loss.detach()), "logits", out.shape)
    siam_goldens()
    vit_goldens()
    vit_mlp_goldens()
    vit_decoder_goldens()
    changeformer_goldens()
    upernet_goldens()


def siam_goldens():
    """FC-Siam-conc / FC-Siam-diff fixtures: train-mode forward+loss+gradients with Dropout2d p=0.2 under a fixed torch
    seed (masks re-derived by oracle.siam_oracle.draw_masks_like_torch and stored), and the eval-mode forward after it."""
    sys.path.insert(0, REF)
    from models.siam_conc import SiamUnet_conc as RefConc     # noqa: E402  (reference, read-only)
    from models.siam_diff import SiamUnet_diff as RefDiff     # noqa: E402
    from utilities.bce_and_dice import BCEandDiceLoss as RefLoss  # noqa: E402
    from oracle import siam_oracle
    from oracle.weights import make_batch

    for kind, Ref, (N, H, W, seed, do_seed) in (("conc", RefConc, (2, 32, 32, 21, 5)), ("diff", RefDiff, (2, 48, 32, 22, 6)),
                                                ("conc", RefConc, (4, 64, 64, 23, 7))):
        sd = siam_oracle.make_state(seed, 2, 3, kind)
        x1, x2, mask = make_batch(seed, N, H, W)
        model = Ref(2, 3)
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
        model.train()
        crit = RefLoss(weights=torch.tensor([1.0, 1.0, 1.0]), ignore_index=3, use_softmax=True)
        masks = siam_oracle.draw_masks_like_torch(do_seed, N, 0.2)
        torch.manual_seed(do_seed)
        out = model(torch.from_numpy(x1), torch.from_numpy(x2))
        loss = crit(out, torch.from_numpy(mask))
        loss.backward()
        # the oracle with the re-derived masks must reproduce the reference (pins both the restatement and the masks)
        osd = siam_oracle.to_torch_state(sd)
        loss_o, out_o, grads_o = siam_oracle.train_step(osd, torch.from_numpy(x1), torch.from_numpy(x2), torch.from_numpy(mask), kind, masks=masks)
        assert torch.allclose(out_o, out.detach(), rtol=1e-4, atol=1e-6), (kind, (out_o - out.detach()).abs().max())
        assert abs(float(loss_o) - float(loss.detach())) < 1e-5
        fx = {"kind": kind, "N": N, "H": H, "W": W, "seed": seed, "out": out.detach().numpy(), "loss": loss.detach().numpy()}
        for k, m in masks.items():
            fx[f"mask.{k}"] = m.numpy()
        names, norms = [], []
        keep_full = {"conv11.weight", "conv11.bias", "bn11.weight", "bn12.bias", "conv22.weight", "conv31.weight", "upconv3.weight", "upconv4.bias",
                     "upconv1.weight", "conv33d.weight", "conv12d.weight", "conv11d.weight", "conv11d.bias", "bn33d.weight", "conv31d.weight"}
        for k, p in model.named_parameters():
            names.append(k)
            norms.append(float(p.grad.double().norm()))
            if k in keep_full:
                fx[f"grad.{k}"] = p.grad.numpy()
        fx["grad_names"] = np.array(names)
        fx["grad_norms"] = np.array(norms, np.float64)
        st = model.state_dict()
        for k in ("bn11.running_mean", "bn11.running_var", "bn11.num_batches_tracked", "bn43.running_var", "bn12d.running_mean",
                  "bn12d.num_batches_tracked"):
            fx[f"state.{k}"] = st[k].numpy()
        model.eval()
        with torch.no_grad():
            fx["out_eval"] = model(torch.from_numpy(x1), torch.from_numpy(x2)).numpy()
        np.savez_compressed(OUT / f"siam_{kind}_n{N}_s{H}x{W}.npz", **fx)
        print("siam", kind, N, H, W, "loss", float(loss.detach()))


def vit_goldens():
    """FloodViT fixtures (ViT encoder + linear FinetunerSegmentation head) from the unmodified reference modules."""
    from oracle.ref_import import install_stubs
    install_stubs()
    from models.vision_transformer import ViT as RefViT                    # noqa: E402  (reference, read-only)
    from models.model_utilities import FinetunerSegmentation as RefFinetuner  # noqa: E402
    from utilities.bce_and_dice import BCEandDiceLoss as RefLoss           # noqa: E402
    from oracle import vit_oracle

    for tag, (dim, depth, heads, mlp, N, seed) in {"d128_l2_h2": (128, 2, 2, 256, 2, 61), "d192_l3_h3": (192, 3, 3, 384, 2, 62)}.items():
        sd = vit_oracle.make_state(seed, dim, depth, heads, mlp)
        img, mask = vit_oracle.make_batch(seed, N)
        enc = RefViT(image_size=224, patch_size=16, num_classes=3, dim=dim, depth=depth, heads=heads, mlp_dim=mlp, channels=6)
        model = RefFinetuner(encoder=enc, configs={"mlp": False, "decoder": False, "num_classes": 3, "finetuning_patch_size": 16})
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
        model.train()
        crit = RefLoss(weights=torch.tensor([1.0, 1.0, 1.0]), ignore_index=3, use_softmax=True)
        out = model(torch.from_numpy(img))
        loss = crit(out, torch.from_numpy(mask))
        loss.backward()
        with torch.no_grad():
            tokens = model.model(torch.from_numpy(img))
        fx = {"dim": dim, "depth": depth, "heads": heads, "mlp": mlp, "N": N, "seed": seed, "loss": loss.detach().numpy(),
              "logits_sample": out.detach().numpy()[:, :, ::7, ::7].copy(),
              "tokens": tokens.numpy()}
        names, norms = [], []
        keep_full = {"model.cls_token", "model.pos_embedding", "model.to_patch_embedding.1.weight", "model.to_patch_embedding.2.bias",
                     "model.to_patch_embedding.3.weight", "model.transformer.norm.bias", "model.transformer.layers.0.0.norm.weight",
                     "model.transformer.layers.0.0.to_qkv.weight", "model.transformer.layers.1.0.to_out.0.weight",
                     "model.transformer.layers.1.1.net.1.bias", "model.transformer.layers.0.1.net.4.weight", "head.weight", "head.bias"}
        for k, p in model.named_parameters():
            names.append(k)
            norms.append(float(p.grad.double().norm()))
            if k in keep_full:
                fx[f"grad.{k}"] = p.grad.numpy()
        fx["grad_names"] = np.array(names)
        fx["grad_norms"] = np.array(norms, np.float64)
        np.savez_compressed(OUT / f"floodvit_{tag}.npz", **fx)
        print("floodvit", tag, "loss", float(loss.detach()))


def vit_mlp_goldens():
    """FloodViT with the `mlp` head (FinetunerSegmentation configs mlp=True) from the unmodified reference modules."""
    from oracle.ref_import import install_stubs
    install_stubs()
    from models.vision_transformer import ViT as RefViT                    # noqa: E402  (reference, read-only)
    from models.model_utilities import FinetunerSegmentation as RefFinetuner  # noqa: E402
    from utilities.bce_and_dice import BCEandDiceLoss as RefLoss           # noqa: E402
    from oracle import vit_oracle
    dim, depth, heads, mlp, N, seed = 128, 2, 2, 256, 2, 63
    sd = vit_oracle.make_state_mlp(seed, dim, depth, heads, mlp)
    img, mask = vit_oracle.make_batch(seed, N)
    enc = RefViT(image_size=224, patch_size=16, num_classes=3, dim=dim, depth=depth, heads=heads, mlp_dim=mlp, channels=6)
    model = RefFinetuner(encoder=enc, configs={"mlp": True, "decoder": False, "num_classes": 3, "finetuning_patch_size": 16})
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    model.train()
    crit = RefLoss(weights=torch.tensor([1.0, 1.0, 1.0]), ignore_index=3, use_softmax=True)
    out = model(torch.from_numpy(img))
    loss = crit(out, torch.from_numpy(mask))
    loss.backward()
    fx = {"dim": dim, "depth": depth, "heads": heads, "mlp": mlp, "N": N, "seed": seed, "loss": loss.detach().numpy(),
          "logits_sample": out.detach().numpy()[:, :, ::7, ::7].copy()}
    names, norms = [], []
    keep_full = {"head.0.weight", "head.0.bias", "head.2.weight", "head.2.bias", "model.transformer.norm.weight",
                 "model.transformer.layers.1.1.net.4.weight", "model.pos_embedding"}
    for k, p in model.named_parameters():
        names.append(k)
        norms.append(float(p.grad.double().norm()))
        if k in keep_full:
            fx[f"grad.{k}"] = p.grad.numpy()
    fx["grad_names"] = np.array(names)
    fx["grad_norms"] = np.array(norms, np.float64)
    np.savez_compressed(OUT / "floodvit_mlp_d128_l2.npz", **fx)
    print("floodvit mlp head loss", float(loss.detach()))


def vit_decoder_goldens():
    """FloodViT with the deconvolution `Decoder` head (FinetunerSegmentation configs decoder=True; encoder width 1024 as the reference
    hard-wires) from the unmodified reference modules."""
    from oracle.ref_import import install_stubs
    install_stubs()
    from models.vision_transformer import ViT as RefViT                    # noqa: E402  (reference, read-only)
    from models.model_utilities import FinetunerSegmentation as RefFinetuner  # noqa: E402
    from utilities.bce_and_dice import BCEandDiceLoss as RefLoss           # noqa: E402
    from oracle import vit_oracle
    depth, heads, mlp, N, seed = 1, 2, 128, 2, 64
    sd = vit_oracle.make_state_decoder(seed, depth, heads, mlp)
    img, mask = vit_oracle.make_batch(seed, N)
    enc = RefViT(image_size=224, patch_size=16, num_classes=3, dim=1024, depth=depth, heads=heads, mlp_dim=mlp, channels=6)
    model = RefFinetuner(encoder=enc, configs={"mlp": False, "decoder": True, "num_classes": 3, "finetuning_patch_size": 16})
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    model.train()
    crit = RefLoss(weights=torch.tensor([1.0, 1.0, 1.0]), ignore_index=3, use_softmax=True)
    out = model(torch.from_numpy(img))
    loss = crit(out, torch.from_numpy(mask))
    loss.backward()
    fx = {"depth": depth, "heads": heads, "mlp": mlp, "N": N, "seed": seed, "loss": loss.detach().numpy(),
          "logits_sample": out.detach().numpy()[:, :, ::7, ::7].copy()}
    names, norms = [], []
    keep_full = {"head.deconv1.bias", "head.deconv2.weight", "head.deconv2.bias", "head.deconv3.weight", "head.deconv3.bias",
                 "model.transformer.norm.weight"}
    for k, p in model.named_parameters():
        names.append(k)
        norms.append(float(p.grad.double().norm()))
        if k in keep_full:
            fx[f"grad.{k}"] = p.grad.numpy()
    fx["grad_names"] = np.array(names)
    fx["grad_norms"] = np.array(norms, np.float64)
    np.savez_compressed(OUT / "floodvit_decoder_d1024_l1.npz", **fx)
    print("floodvit decoder head loss", float(loss.detach()))


def changeformer_goldens():
    """ChangeFormerV6 fixtures from the unmodified reference (timm stubbed: DropPath / trunc_normal_ / to_2tuple only); the stochastic
    layers run with p = 0 (train-mode BatchNorm statistics kept), see oracle/changeformer_oracle.py."""
    from oracle.ref_import import install_stubs
    install_stubs()
    from models.changeformer import ChangeFormerV6 as RefCF                # noqa: E402  (reference, read-only)
    from utilities.bce_and_dice import BCEandDiceLoss as RefLoss           # noqa: E402
    from oracle import changeformer_oracle as co
    from oracle.weights import make_batch

    for tag, (N, HW, seed) in {"n2_s64": (2, 64, 81), "n2_s224": (2, 224, 82)}.items():
        sd = co.make_state(seed)
        x1, x2, mask = make_batch(seed, N, HW, HW)
        model = RefCF(embed_dim=256, input_nc=2, output_nc=3, decoder_softmax=True)
        assert list(model.state_dict().keys()) == list(sd.keys())
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if type(mod).__name__ == "DropPath":
                mod.drop_prob = 0.0
        model.train()
        crit = RefLoss(weights=torch.tensor([1.0, 1.0, 1.0]), ignore_index=3, use_softmax=True)
        outs = model(torch.from_numpy(x1), torch.from_numpy(x2))
        loss = crit(outs[-1], torch.from_numpy(mask))          # change_detection_trainer.py:166-170 (no multi_scale_train)
        loss.backward()
        st = max(1, HW // 32)
        fx = {"N": N, "HW": HW, "seed": seed, "loss": loss.detach().numpy(), "out_sample": outs[-1].detach().numpy()[:, :, ::st, ::st].copy()}
        for i in range(4):
            fx[f"side{i}"] = outs[i].detach().numpy()
        names, norms = [], []
        keep_full = {"Tenc_x2.patch_embed1.proj.weight", "Tenc_x2.patch_embed1.norm.bias", "Tenc_x2.block1.0.attn.sr.weight",
                     "Tenc_x2.block1.0.attn.q.bias", "Tenc_x2.block2.1.attn.kv.weight", "Tenc_x2.block1.2.mlp.dwconv.dwconv.weight",
                     "Tenc_x2.block3.0.norm1.weight", "Tenc_x2.block4.2.mlp.fc2.bias", "Tenc_x2.norm4.weight", "TDec_x2.linear_c1.proj.weight",
                     "TDec_x2.diff_c3.2.weight", "TDec_x2.diff_c4.0.bias", "TDec_x2.linear_fuse.1.bias", "TDec_x2.convd1x.conv2d.bias",
                     "TDec_x2.change_probability.conv2d.weight", "TDec_x2.change_probability.conv2d.bias"}
        for k, p in model.named_parameters():
            names.append(k)
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            norms.append(float(g.double().norm()))
            if k in keep_full:
                fx[f"grad.{k}"] = g.numpy()
        fx["grad_names"] = np.array(names)
        fx["grad_norms"] = np.array(norms, np.float64)
        stt = model.state_dict()
        for k in ("TDec_x2.diff_c1.2.running_mean", "TDec_x2.diff_c4.2.running_var", "TDec_x2.linear_fuse.1.running_mean",
                  "TDec_x2.make_pred_c2.2.running_var", "TDec_x2.linear_fuse.1.num_batches_tracked"):
            fx[f"state.{k}"] = stt[k].numpy()
        model.eval()
        with torch.no_grad():
            fx["out_eval_sample"] = model(torch.from_numpy(x1), torch.from_numpy(x2))[-1].numpy()[:, :, ::st, ::st].copy()
        np.savez_compressed(OUT / f"changeformer_{tag}.npz", **fx)
        print("changeformer", tag, "loss", float(loss.detach()))


def upernet_goldens():
    """FloodViT + UPerNet: the reference's ViT modules (driven layer by layer to tap the residual stream) + the installed HF
    UperNetHead class (the third-party code models/upernet.py:80 reaches), logits resized like UperNetForSemanticSegmentation."""
    from oracle.ref_import import install_stubs
    install_stubs()
    from models.vision_transformer import ViT as RefViT                    # noqa: E402  (reference, read-only)
    from utilities.bce_and_dice import BCEandDiceLoss as RefLoss           # noqa: E402
    from transformers import UperNetConfig
    from transformers.models.upernet.modeling_upernet import UperNetHead
    from einops import repeat
    from oracle import upernet_oracle as uo
    from oracle import vit_oracle
    import torch.nn.functional as F

    dim, depth, heads, mlp, N, seed, out_idx = 128, 4, 2, 256, 8, 71, [1, 2, 3, 4]   # N=8: the pool-scale-1 module normalises over N samples; N=2 makes its 1/sqrt(var+eps) reach 316 and the case ill-conditioned
    sd = uo.make_state(seed, dim, depth, heads, mlp)
    img, mask = vit_oracle.make_batch(seed, N)
    enc = RefViT(image_size=224, patch_size=16, num_classes=3, dim=dim, depth=depth, heads=heads, mlp_dim=mlp, channels=6)
    enc.mlp_head = torch.nn.Identity()
    head = UperNetHead(UperNetConfig(hidden_size=512, pool_scales=[1, 2, 3, 6], num_labels=3), in_channels=[dim] * 4)
    enc.load_state_dict({k[len("model."):]: torch.from_numpy(np.array(v)) for k, v in sd.items() if k.startswith("model.")})
    head.load_state_dict({k[len("decode_head."):]: torch.from_numpy(np.array(v)) for k, v in sd.items() if k.startswith("decode_head.")})
    enc.train(); head.train()
    x = enc.to_patch_embedding(torch.from_numpy(img))
    b, n, _ = x.shape
    x = torch.cat((repeat(enc.cls_token, "1 1 d -> b 1 d", b=b), x), dim=1) + enc.pos_embedding[:, : n + 1]
    taps = {}
    for l, (attn, ff) in enumerate(enc.transformer.layers):
        x = attn(x) + x
        x = ff(x) + x
        taps[l + 1] = x[:, 1:]
    taps[depth] = enc.transformer.norm(x)[:, 1:]
    feats = [taps[i].reshape(b, 14, 14, dim).permute(0, 3, 1, 2) for i in out_idx]
    logits = F.interpolate(head(feats), size=(224, 224), mode="bilinear", align_corners=False)
    crit = RefLoss(weights=torch.tensor([1.0, 1.0, 1.0]), ignore_index=3, use_softmax=True)
    loss = crit(logits, torch.from_numpy(mask))
    loss.backward()
    fx = {"dim": dim, "depth": depth, "heads": heads, "mlp": mlp, "N": N, "seed": seed, "out_indices": np.array(out_idx),
          "loss": loss.detach().numpy(), "logits_sample": logits.detach().numpy()[:, :, ::7, ::7].copy()}
    names, norms = [], []
    keep_full = {"model.cls_token", "model.transformer.layers.0.0.to_qkv.weight", "model.transformer.layers.2.1.net.4.weight",
                 "decode_head.classifier.weight", "decode_head.psp_modules.0.1.conv.weight", "decode_head.psp_modules.2.1.batch_norm.weight",
                 "decode_head.lateral_convs.1.conv.weight", "decode_head.fpn_convs.0.batch_norm.bias", "decode_head.fpn_bottleneck.batch_norm.weight"}
    for pre, mod in (("model.", enc), ("decode_head.", head)):
        for k, p in mod.named_parameters():
            names.append(pre + k)
            norms.append(float(p.grad.double().norm()))
            if pre + k in keep_full:
                fx[f"grad.{pre + k}"] = p.grad.numpy()
    fx["grad_names"] = np.array(names)
    fx["grad_norms"] = np.array(norms, np.float64)
    for k in ("bottleneck.batch_norm.running_mean", "psp_modules.3.1.batch_norm.running_var", "fpn_bottleneck.batch_norm.running_var"):
        fx[f"state.decode_head.{k}"] = head.state_dict()[k].numpy()
    np.savez_compressed(OUT / "floodvit_upernet_d128_l4.npz", **fx)
    print("floodvit+upernet loss", float(loss.detach()))


def bf16_drift_goldens():
    """How far the UNMODIFIED reference moves when it runs under torch.autocast(bfloat16) instead of fp32 - per gradient tensor, on
    exactly the cases the bf16 GPU tests use.  The tests bound the CUDA bf16 path's gradient error by a multiple of THIS drift
    (tests/golden/bf16_drift.npz) instead of by a free constant: the bf16 bar is pinned to the reference's own behaviour."""
    from oracle.ref_import import install_stubs
    install_stubs()
    sys.path.insert(0, REF)
    from models.changeformer import ChangeFormerV6 as RefCF
    from models.siam_conc import SiamUnet_conc as RefConc
    from models.snunet import SNUNet_ECAM as RefSNUNet
    from utilities.bce_and_dice import BCEandDiceLoss as RefLoss
    from oracle import changeformer_oracle as co
    from oracle import siam_oracle
    from oracle.weights import make_batch, make_state
    torch.set_num_threads(8)
    out = {}

    def run(tag, build, sd, batch, pick_last=False):
        grads = {}
        for mode in ("fp32", "bf16"):
            model = build()
            model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
            for mod in model.modules():
                if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
                    mod.p = 0.0
                if type(mod).__name__ == "DropPath":
                    mod.drop_prob = 0.0
            if hasattr(model, "do11"):                       # FC-Siam: functional dropout2d objects
                for k_, v_ in vars(model)["_modules"].items():
                    if k_.startswith("do"):
                        v_.p = 0.0
            model.train()
            crit = RefLoss(weights=torch.tensor([1.0, 1.0, 1.0]), ignore_index=3, use_softmax=True)
            x1, x2, mask = (torch.from_numpy(a) for a in batch)
            with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                o = model(x1, x2)
                o = o[-1] if pick_last else o
            loss = crit(o.float(), mask)
            loss.backward()
            grads[mode] = {k: p.grad.detach().double().clone() for k, p in model.named_parameters() if p.grad is not None}
            out[f"{tag}.loss.{mode}"] = float(loss.detach())
            out[f"{tag}.out_rel_l2"] = 0.0
            if mode == "fp32":
                ref_out = o.detach().double()
            else:
                out[f"{tag}.out_rel_l2"] = float((o.detach().double() - ref_out).norm() / ref_out.norm())
        names, drift = [], []
        for k, g in grads["fp32"].items():
            names.append(k)
            drift.append(float((grads["bf16"][k] - g).norm() / (g.norm() + 1e-300)))
        out[f"{tag}.names"] = np.array(names)
        out[f"{tag}.drift"] = np.array(drift, np.float64)
        print(tag, "out rel-L2", out[f"{tag}.out_rel_l2"], "max drift", max(drift), "median", float(np.median(drift)))

    run("snunet_b32_n2_s64_seed21", lambda: RefSNUNet(2, 3, base_channel=32), make_state(21, 2, 3, 32), make_batch(21, 2, 64, 64))
    from models.siam_diff import SiamUnet_diff as RefDiff
    run("siam_conc_n2_s64_seed33", lambda: RefConc(2, 3), siam_oracle.make_state(33, 2, 3, "conc"), make_batch(33, 2, 64, 64))
    run("siam_diff_n2_s64_seed33", lambda: RefDiff(2, 3), siam_oracle.make_state(33, 2, 3, "diff"), make_batch(33, 2, 64, 64))
    run("changeformer_n2_s224_seed91", lambda: RefCF(embed_dim=256, input_nc=2, output_nc=3, decoder_softmax=True), co.make_state(91),
        make_batch(91, 2, 224, 224), pick_last=True)
    # ViT-B/16 (768 / 12 / 12 / 3072, 6 channels: the BASELINE.json encoder), N = 2: single-input model
    from models.model_utilities import FinetunerSegmentation as RefFinetuner
    from models.vision_transformer import ViT as RefViT
    from oracle import vit_oracle

    class _Seg(torch.nn.Module):                              # adapter: run(...) feeds (x1, x2); the ViT takes one stacked image
        def __init__(self):
            super().__init__()
            enc = RefViT(image_size=224, patch_size=16, num_classes=3, dim=768, depth=12, heads=12, mlp_dim=3072, channels=6)
            self.m = RefFinetuner(encoder=enc, configs={"mlp": False, "decoder": False, "num_classes": 3, "finetuning_patch_size": 16})

        def forward(self, img, _unused):
            return self.m(img)

        def load_state_dict(self, sd, **kw):
            return self.m.load_state_dict(sd, **kw)

        def named_parameters(self, *a, **k):
            return self.m.named_parameters(*a, **k)
    img, mask = vit_oracle.make_batch(71, 2)
    run("floodvit_b_n2_seed71", _Seg, vit_oracle.make_state(71, 768, 12, 12, 3072), (img, img[:, :1], mask))
    np.savez_compressed(OUT / "bf16_drift.npz", **out)


if __name__ == "__main__":
    if "--bf16-drift-only" in sys.argv:
        bf16_drift_goldens()
        sys.exit(0)
    if "--upernet-only" in sys.argv:
        upernet_goldens()
        sys.exit(0)
    if "--cf-only" in sys.argv:
        changeformer_goldens()
        sys.exit(0)
    if "--vit-only" in sys.argv:
        vit_goldens()
        sys.exit(0)
    if "--vit-decoder-only" in sys.argv:
        vit_decoder_goldens()
        sys.exit(0)
    if "--vit-mlp-only" in sys.argv:
        vit_mlp_goldens()
        sys.exit(0)
    if "--siam-only" in sys.argv:
        siam_goldens()
        sys.exit(0)
    main()
