"""Files written BY THE REFERENCE load into the B200 mirrors: the change-detection checkpoint dict
(reference training/change_detection_trainer.py:312-318) and the pickled segmentation module (segmentation_trainer.py:255,
main.py:151; models/model_utilities.py:159 for the encoder).  Needs the reference sources (build container); skipped elsewhere."""
import importlib
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference sources not present (GPU box)")


def _ref(module):
    from oracle.ref_import import install_stubs
    install_stubs()
    sys.modules.pop("models", None)
    return importlib.import_module(f"models.{module}")


def test_reference_cd_checkpoint_dict_loads(tmp_path):
    from kurosiwo_b200.checkpoint_compat import load_reference_checkpoint
    from kurosiwo_b200.snunet import SNUNet_ECAM
    torch.manual_seed(3)
    ref = _ref("snunet").SNUNet_ECAM(2, 3, base_channel=8)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, 4)
    torch.save({"epoch": 2, "model_state_dict": ref.state_dict(), "optimizer_state_dict": opt.state_dict(),
                "lr_scheduler_state_dict": sched.state_dict(), "loss": 0.5}, tmp_path / "best_segmentation.pt")
    ours = SNUNet_ECAM(2, 3, base_channel=8)
    out = load_reference_checkpoint(tmp_path / "best_segmentation.pt", ours)
    assert out is ours
    for k, v in ref.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k


@pytest.mark.parametrize("head", ["linear", "mlp"])
def test_reference_pickled_segmentation_module_converts(tmp_path, head):
    from kurosiwo_b200.checkpoint_compat import load_reference_checkpoint
    from kurosiwo_b200.vision_transformer import FinetunerSegmentation
    torch.manual_seed(4)
    enc = _ref("vision_transformer").ViT(image_size=224, patch_size=16, num_classes=10, dim=128, depth=2, heads=2, mlp_dim=256, channels=6)
    cfg = {"mlp": head == "mlp", "decoder": False, "num_classes": 3, "finetuning_patch_size": 16, "image_size": 224}
    ref = _ref("model_utilities").FinetunerSegmentation(encoder=enc, configs=cfg)
    torch.save(ref, tmp_path / "best_segmentation.pt")                 # what segmentation_trainer.py:255 writes
    ours = load_reference_checkpoint(tmp_path / "best_segmentation.pt", None, "cpu", "fp32")
    assert isinstance(ours, FinetunerSegmentation) and type(ours).__module__.startswith("kurosiwo_b200")
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert [k for k in rs] == [k for k in os_ if not k.startswith("model.mlp_head.")]
    for k, v in rs.items():
        assert torch.equal(v, os_[k]), k
    # the encoder alone, as models/model_utilities.py:159 loads it
    torch.save(enc, tmp_path / "encoder.pt")
    from kurosiwo_b200.checkpoint_compat import from_reference_module
    e2 = from_reference_module(torch.load(tmp_path / "encoder.pt", weights_only=False), "fp32")
    assert e2.cfg["dim"] == 128 and e2.cfg["depth"] == 2 and e2.cfg["heads"] == 2 and e2.cfg["channels"] == 6
