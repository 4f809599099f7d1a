"""Same seed + construct == the reference's construct, for every model class on the path (the constructors
register parameters in the reference order and draw from the same RNG stream, so `torch.manual_seed(s); Model(...)` gives a
state dict that is BIT-identical to the reference class's).  Needs the reference sources (build container); skipped elsewhere.
Test infrastructure: imports the UNMODIFIED reference through oracle/ref_import.py, never from the product."""
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
    sys.modules.pop("models", None)            # the repo has no `models` package; make sure the reference's namespace package wins
    return importlib.import_module(f"models.{module}")


def _same(ours, ref):
    a, b = ours.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("seed", [0, 999])
def test_snunet_init(seed):
    from kurosiwo_b200.snunet import SNUNet_ECAM
    torch.manual_seed(seed); ours = SNUNet_ECAM(2, 3, base_channel=32)
    torch.manual_seed(seed); ref = _ref("snunet").SNUNet_ECAM(2, 3, base_channel=32)
    _same(ours, ref)


@pytest.mark.parametrize("kind", ["conc", "diff"])
def test_siam_init(kind):
    from kurosiwo_b200 import siam_unet
    ours_cls = getattr(siam_unet, f"SiamUnet_{kind}")
    ref_cls = getattr(_ref(f"siam_{kind}"), f"SiamUnet_{kind}")
    torch.manual_seed(5); ours = ours_cls(input_nbr=2, label_nbr=3)
    torch.manual_seed(5); ref = ref_cls(2, 3)
    _same(ours, ref)


def test_changeformer_init():
    from kurosiwo_b200.changeformer import ChangeFormerV6
    torch.manual_seed(7); ours = ChangeFormerV6(input_nc=2, output_nc=3, decoder_softmax=True, embed_dim=256)
    torch.manual_seed(7); ref = _ref("changeformer").ChangeFormerV6(input_nc=2, output_nc=3, decoder_softmax=True, embed_dim=256)
    _same(ours, ref)


def test_vit_init():
    from kurosiwo_b200.vision_transformer import ViT
    kw = dict(image_size=224, patch_size=16, num_classes=3, dim=192, depth=3, heads=3, mlp_dim=384, channels=6)
    torch.manual_seed(11); ours = ViT(**kw)
    torch.manual_seed(11); ref = _ref("vision_transformer").ViT(**kw)
    _same(ours, ref)
