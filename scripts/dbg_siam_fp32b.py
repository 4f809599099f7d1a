import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from pathlib import Path
from kurosiwo_b200.siam_unet import SiamUnet_conc
from kurosiwo_b200.bce_and_dice import BCEandDiceLoss
from oracle import siam_oracle, weights
from oracle.snunet_oracle import ce_dice_torch
DEV = "cuda:0"
fx = np.load(Path('tests/golden') / 'siam_conc_n4_s64x64.npz')
kind, N, H, W, seed = str(fx["kind"]), int(fx["N"]), int(fx["H"]), int(fx["W"]), int(fx["seed"])
sd_np = siam_oracle.make_state(seed, 2, 3, kind)
x1, x2, mask = (torch.from_numpy(a) for a in weights.make_batch(seed, N, H, W))
sd = siam_oracle.to_torch_state(sd_np)
leaves = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("running_mean", "running_var"))) for k, v in sd.items()}
tap = {}
out_o = siam_oracle.siam_forward(leaves, x1, x2, kind, True, None, tap)
for t in tap.values(): t.retain_grad()
loss_o = ce_dice_torch(out_o, mask, (1., 1., 1.)); loss_o.backward()
m = SiamUnet_conc(2, 3, precision="fp32")
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
m = m.to(DEV).train(); m.dropout_p = 0.0
crit = BCEandDiceLoss(weights=[1.0, 1.0, 1.0], ignore_index=3, use_softmax=True).to(DEV)
out = m(x1.to(DEV), x2.to(DEV)); loss = crit(out, mask.to(DEV)); loss.backward()
eng = m.engine(x1.to(DEV))
for L in eng.layers:
    key = L.mask_key
    a = L.out.v.tensor().float().cpu().permute(0, 3, 1, 2); ao = tap[key].detach()
    g = L.out.g.tensor().float().cpu().permute(0, 3, 1, 2); go = tap[key].grad
    # NOTE: L.out.g holds the MASKED gradient after backward (bn_bwd_reduce masks in place): compare on out>0
    go_m = go * (ao > 0)
    print(f"{key:6s} act err {((a-ao).abs().max()/ao.abs().max()).item():.2e}  grad err {((g-go_m).abs().max()/go_m.abs().max()).item():.2e}  n_bad {(((g-go_m).abs() > 1e-3*go_m.abs().max()).sum()).item()}")
