"""Checkpoint compatibility with the reference.

The reference writes two kinds of files:
  * change detection: `torch.save({'epoch', 'model_state_dict', 'optimizer_state_dict', 'lr_scheduler_state_dict', 'loss'})`
    (training/change_detection_trainer.py:206-213, :312-318) - `model_state_dict` has the reference key names, which the mirrors in
    this package reproduce one to one, so `load_state_dict` just works;
  * segmentation / FloodViT: the WHOLE module pickled with `torch.save(model, ...)` (training/segmentation_trainer.py:255; the published
    FloodViT encoder is loaded the same way, models/model_utilities.py:159).  Unpickling such a file needs the reference's classes
    importable (its repository on PYTHONPATH); what comes out is a reference `nn.Module`, which `from_reference_module` converts into
    the B200 mirror of the same architecture: hyper-parameters are read off the state-dict shapes, weights are copied by name.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


def _vit_kwargs(sd, prefix=""):
    pos, qkv = sd[f"{prefix}pos_embedding"], sd[f"{prefix}transformer.layers.0.0.to_qkv.weight"]
    dim = pos.shape[2]
    depth = 1 + max(int(k[len(prefix):].split(".")[2]) for k in sd if k.startswith(f"{prefix}transformer.layers."))
    patch_dim = sd[f"{prefix}to_patch_embedding.1.weight"].shape[0]
    n_patches = pos.shape[1] - 1
    side = int(round(n_patches ** 0.5))
    patch = 224 // side
    out = dict(image_size=side * patch, patch_size=patch, dim=dim, depth=depth, heads=qkv.shape[0] // (3 * 64), dim_head=64,
               mlp_dim=sd[f"{prefix}transformer.layers.0.1.net.1.weight"].shape[0], channels=patch_dim // (patch * patch))
    mh = sd.get(f"{prefix}mlp_head.weight")
    out["num_classes"] = mh.shape[0] if mh is not None else 1000
    return out


def from_reference_module(ref: nn.Module, precision: str = "bf16", configs: Optional[dict] = None) -> nn.Module:
    """A reference nn.Module (or one of this package's) -> the kurosiwo_b200 module of the same architecture with the same weights."""
    from .changeformer import ChangeFormerV6
    from .siam_unet import SiamUnet_conc, SiamUnet_diff
    from .snunet import SNUNet_ECAM
    from .vision_transformer import FinetunerSegmentation, ViT
    if type(ref).__module__.startswith("kurosiwo_b200"):
        return ref
    sd = ref.state_dict()
    name = type(ref).__name__
    if name == "SNUNet_ECAM":
        w = sd["conv0_0.conv1.weight"]
        new = SNUNet_ECAM(w.shape[1], sd["conv_final.weight"].shape[0], base_channel=w.shape[0], precision=precision)
    elif name in ("SiamUnet_conc", "SiamUnet_diff"):
        cls = SiamUnet_conc if name.endswith("conc") else SiamUnet_diff
        new = cls(input_nbr=sd["conv11.weight"].shape[1], label_nbr=sd["conv11d.weight"].shape[1], precision=precision)
    elif name == "ChangeFormerV6":
        new = ChangeFormerV6(input_nc=sd["Tenc_x2.patch_embed1.proj.weight"].shape[1], output_nc=sd["TDec_x2.change_probability.conv2d.weight"].shape[0],
                             decoder_softmax=True, embed_dim=sd["TDec_x2.linear_fuse.0.weight"].shape[0], precision=precision)
    elif name == "ViT":
        kw = _vit_kwargs(sd)
        new = ViT(precision=precision, **kw)
    elif name == "FinetunerSegmentation":
        kw = _vit_kwargs(sd, "model.")
        kw.pop("num_classes")
        enc = ViT(num_classes=1000, precision=precision, **kw)
        cfg = dict(configs or getattr(ref, "configs", None) or {})
        cfg.setdefault("finetuning_patch_size", kw["patch_size"])
        cfg["mlp"] = "head.2.weight" in sd and "head.0.weight" in sd and sd["head.0.weight"].dim() == 4 and "head.4.weight" not in sd
        cfg["decoder"] = any(k.startswith("head.") and "." in k[5:] and not k[5].isdigit() for k in sd)
        cfg["num_classes"] = cfg.get("num_classes", 3)
        new = FinetunerSegmentation(encoder=enc, configs=cfg, precision=precision)
        sd = {k: v for k, v in sd.items() if not k.startswith("model.mlp_head.")}   # the reference replaces mlp_head by nn.Identity
    else:
        raise NotImplementedError(f"no B200 mirror for reference module {type(ref).__module__}.{name} (SURVEY.md section 8 scope)")
    missing, unexpected = new.load_state_dict(sd, strict=False)
    missing = [k for k in missing if not k.startswith(("model.mlp_head.", "mlp_head."))]     # FinetunerSegmentation sets mlp_head = nn.Identity()
    if missing or unexpected:
        raise RuntimeError(f"state-dict mismatch converting {name}: missing {missing[:5]}, unexpected {list(unexpected)[:5]}")
    return new


def load_reference_checkpoint(path, model: Optional[nn.Module] = None, device="cpu", precision: str = "bf16", configs: Optional[dict] = None):
    """Load either kind of reference checkpoint.  dict -> `model.load_state_dict(ckpt['model_state_dict'])` (model required);
    pickled module -> converted with from_reference_module.  Returns the (kurosiwo_b200) module."""
    ckpt = torch.load(path, map_location=device, weights_only=False)
    if isinstance(ckpt, nn.Module):
        return from_reference_module(ckpt, precision, configs).to(device)
    if model is None:
        raise ValueError("a state-dict checkpoint needs the model to load into")
    model.load_state_dict(ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt)
    return model
