"""Mirror of the reference's `models/model_utilities.py:initialize_cd_model` for the models on the B200 path."""
from __future__ import annotations

import torch

from .changeformer import ChangeFormerV6
from .siam_unet import SiamUnet_conc, SiamUnet_diff
from .snunet import SNUNet_ECAM
from .vision_transformer import FinetunerSegmentation, FloodViTUperNet, ViT  # noqa: F401


def initialize_cd_model(configs, model_configs, phase="train"):
    method = configs["method"].lower()
    precision = "bf16" if configs.get("mixed_precision", True) else "fp32"
    precision = configs.get("precision", precision)
    if method == "siam-conc":                              # model_utilities.py:183-186
        model = SiamUnet_conc(input_nbr=configs["num_channels"], label_nbr=configs["num_classes"], precision=precision)
    elif method == "siam-diff":                            # model_utilities.py:187-190
        model = SiamUnet_diff(input_nbr=configs["num_channels"], label_nbr=configs["num_classes"], precision=precision)
    elif method == "changeformer":                         # model_utilities.py:199-205
        model = ChangeFormerV6(embed_dim=model_configs["embed_dim"], input_nc=configs["num_channels"], output_nc=configs["num_classes"],
                               decoder_softmax=model_configs["decoder_softmax"], precision=precision)
    elif method == "snunet":
        model = SNUNet_ECAM(configs["num_channels"], configs["num_classes"], base_channel=model_configs["base_channel"],
                            precision=precision)
    else:
        raise NotImplementedError(f"method {configs['method']} is not on the B200 hot path yet (SURVEY.md §8: "
                                  "bit-cd, hfa-net, adhr-cdnet and transunet-cd are outside it)")
    model = model.to(configs["device"])
    if configs.get("resume_checkpoint"):
        checkpoint = torch.load(configs["resume_checkpoint"], map_location=configs["device"])
        model.load_state_dict(checkpoint["model_state_dict"])
    return model


def initialize_segmentation_model(config, model_configs):
    """models/model_utilities.py:97-167 for the method on the B200 path: `finetune` (FloodViT).  The reference loads a pickled
    encoder from config["encoder"] (:159); without a checkpoint the encoder is built from the `encoder_config` block of
    configs/method/finetune/finetune.json (the reference ships neither - SURVEY.md §8(c) "FloodViT definition gap")."""
    if config["method"].lower() != "finetune":
        raise NotImplementedError(f"method {config['method']} is not on the B200 hot path (SURVEY.md §8: smp U-Net family, UPerNet backbones)")
    precision = "bf16" if config.get("mixed_precision", True) else "fp32"
    precision = config.get("precision", precision)
    if config.get("encoder"):
        from .checkpoint_compat import from_reference_module     # a ViT pickled by the reference (its classes on PYTHONPATH) or by us
        encoder = from_reference_module(torch.load(config["encoder"], map_location="cpu", weights_only=False), precision)
        encoder.precision = precision
    else:
        ec = dict(model_configs.get("encoder_config") or config["encoder_config"])
        encoder = ViT(image_size=ec.get("image_size", 224), patch_size=ec.get("patch_size", 16), num_classes=ec.get("num_classes", 1000),
                      dim=ec["dim"], depth=ec["depth"], heads=ec["heads"], mlp_dim=ec["mlp_dim"], channels=config["num_channels"],
                      precision=precision)
    for param in encoder.parameters():
        param.requires_grad = not config.get("linear_eval", False)
    if config.get("head", model_configs.get("head", "linear")) == "upernet":      # BASELINE.json configs[3]: MAE-ViT encoder + UPerNet head
        return FloodViTUperNet(encoder, num_classes=config["num_classes"], hidden_size=int(model_configs.get("upernet_hidden", 512)),
                               out_indices=model_configs.get("out_indices"), precision=precision)
    return FinetunerSegmentation(encoder=encoder, configs=config, precision=precision)
