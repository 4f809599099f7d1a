import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle import upernet_oracle as uo, vit_oracle
from kurosiwo_b200.vision_transformer import FloodViTUperNet, ViT
from kurosiwo_b200.bce_and_dice import BCEandDiceLoss
fx = np.load("tests/golden/floodvit_upernet_d128_l4.npz")
dim, depth, heads, mlp, N, seed = (int(fx[k]) for k in ("dim", "depth", "heads", "mlp", "N", "seed"))
out_idx = [int(i) for i in fx["out_indices"]]
sd_np = uo.make_state(seed, dim, depth, heads, mlp)
img, mask = (torch.from_numpy(a) for a in vit_oracle.make_batch(seed, N))
loss_o, logits_o, grads_o = uo.train_step(vit_oracle.to_torch_state(sd_np), img, mask, heads, out_idx)
enc = ViT(image_size=224, patch_size=16, num_classes=3, dim=dim, depth=depth, heads=heads, mlp_dim=mlp, channels=6, precision="fp32")
m = FloodViTUperNet(enc, num_classes=3, hidden_size=512, out_indices=out_idx)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
m = m.cuda().train()
crit = BCEandDiceLoss(weights=[1.0, 1.0, 1.0], ignore_index=3, use_softmax=True).cuda()
out = m(img.cuda()); loss = crit(out, mask.cuda()); loss.backward()
print("loss", loss.item(), float(loss_o), "logits", float((out.cpu() - logits_o).abs().max()))
for n, p in m.named_parameters():
    go = grads_o[n]; e = float((p.grad.cpu() - go).abs().max()); sc = float(go.abs().max())
    print(f"{n:60s} {e:.3e} {sc:.3e} {e/(sc+1e-12):.3e}")
