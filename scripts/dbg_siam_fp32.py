import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from pathlib import Path
from kurosiwo_b200.siam_unet import SiamUnet_conc
from kurosiwo_b200.bce_and_dice import BCEandDiceLoss
from oracle import siam_oracle, weights
DEV = "cuda:0"
fx = np.load(Path('tests/golden') / 'siam_conc_n4_s64x64.npz')
kind, N, H, W, seed = str(fx["kind"]), int(fx["N"]), int(fx["H"]), int(fx["W"]), int(fx["seed"])
sd_np = siam_oracle.make_state(seed, 2, 3, kind)
x1, x2, mask = (torch.from_numpy(a) for a in weights.make_batch(seed, N, H, W))
masks = {k[5:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("mask.")}
sd = siam_oracle.to_torch_state(sd_np)
loss_o, out_o, grads_o = siam_oracle.train_step(sd, x1, x2, mask, kind, masks=masks)
for use_masks in (True, False):
    m = SiamUnet_conc(2, 3, precision="fp32")
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
    m = m.to(DEV).train()
    if use_masks: m.engine(x1.to(DEV)).fixed_masks = masks
    else:
        m.dropout_p = 0.0
        sd2 = siam_oracle.to_torch_state(sd_np)
        loss_o, out_o, grads_o = siam_oracle.train_step(sd2, x1, x2, mask, kind, masks=None)
    crit = BCEandDiceLoss(weights=[1.0, 1.0, 1.0], ignore_index=3, use_softmax=True).to(DEV)
    out = m(x1.to(DEV), x2.to(DEV)); loss = crit(out, mask.to(DEV)); loss.backward()
    print("masks", use_masks, "out err", ((out.detach().cpu() - out_o).abs().max() / out_o.abs().max()).item(), "loss", loss.item(), float(loss_o))
    for n, p in m.named_parameters():
        go = grads_o[n]; e = (p.grad.cpu() - go).abs().max().item(); s = go.abs().max().item()
        if e > 1e-4 * s + 1e-9 and not (n.endswith('.bias') and n.startswith('conv') and n != 'conv11d.bias'):
            print(f"   {n:18s} err {e:.3e} scale {s:.3e} ratio {e / s:.2e}")
