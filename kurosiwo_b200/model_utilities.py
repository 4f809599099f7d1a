"""Mirror of the reference's `models/model_utilities.py:initialize_cd_model` for the models on the B200 path."""
from __future__ import annotations

import torch

from .siam_unet import SiamUnet_conc, SiamUnet_diff
from .snunet import SNUNet_ECAM


def initialize_cd_model(configs, model_configs, phase="train"):
    method = configs["method"].lower()
    precision = "bf16" if configs.get("mixed_precision", True) else "fp32"
    precision = configs.get("precision", precision)
    if method == "siam-conc":                              # model_utilities.py:183-186
        model = SiamUnet_conc(input_nbr=configs["num_channels"], label_nbr=configs["num_classes"], precision=precision)
    elif method == "siam-diff":                            # model_utilities.py:187-190
        model = SiamUnet_diff(input_nbr=configs["num_channels"], label_nbr=configs["num_classes"], precision=precision)
    elif method == "snunet":
        model = SNUNet_ECAM(configs["num_channels"], configs["num_classes"], base_channel=model_configs["base_channel"],
                            precision=precision)
    else:
        raise NotImplementedError(f"method {configs['method']} is not on the B200 hot path yet (SURVEY.md §8: "
                                  "changeformer and finetune are 'next' rows)")
    model = model.to(configs["device"])
    if configs.get("resume_checkpoint"):
        checkpoint = torch.load(configs["resume_checkpoint"], map_location=configs["device"])
        model.load_state_dict(checkpoint["model_state_dict"])
    return model
