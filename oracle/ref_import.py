"""Import helpers for the UNMODIFIED reference under /root/reference (build container only; TEST INFRASTRUCTURE).

The reference's modules import third-party packages that are not installed here (timm, segmentation_models_pytorch,
denoising_diffusion_pytorch, ...).  None of them is on the arithmetic path of the models we pin, so they are replaced by
inert `sys.modules` stubs; `timm.models.layers` gets the three trivial symbols ChangeFormer needs
(DropPath, to_2tuple, trunc_normal_ - semantics of timm==0.6.12, requirements.txt:9).
"""
from __future__ import annotations

import sys
import types

REF = "/root/reference"


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


def install_stubs():
    import torch
    import torch.nn as nn

    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:                                  # models/upernet.py imports transformers, whose availability probes must run BEFORE
        import transformers               # the timm stub (a module without __spec__) is installed
        from transformers import UperNetConfig  # noqa: F401  (forces the lazy module to resolve its probes now)
    except Exception:
        pass
    for name in ("segmentation_models_pytorch", "denoising_diffusion_pytorch", "pyjson5", "torchmetrics", "kornia", "albumentations",
                 "compress_pickle", "richdem", "rioxarray", "torchio", "torchio.transforms", "vit_pytorch", "vit_pytorch.vit", "torchsummary",
                 "wandb", "kornia.augmentation", "cv2", "rasterio", "geopandas", "shapely", "skimage", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    if "timm" not in sys.modules:
        try:
            import timm  # noqa: F401
        except Exception:
            timm = types.ModuleType("timm")
            models = types.ModuleType("timm.models")
            layers = types.ModuleType("timm.models.layers")

            class DropPath(nn.Module):          # timm 0.6.12 drop_path: per-sample stochastic depth, identity in eval / p == 0
                def __init__(self, drop_prob=0.0, scale_by_keep=True):
                    super().__init__()
                    self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

                def forward(self, x):
                    if self.drop_prob == 0.0 or not self.training:
                        return x
                    keep = 1 - self.drop_prob
                    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
                    r = x.new_empty(shape).bernoulli_(keep)
                    if keep > 0.0 and self.scale_by_keep:
                        r.div_(keep)
                    return x * r

            def to_2tuple(x):
                return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

            def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
                return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

            layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
            timm.models, models.layers = models, layers
            sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
