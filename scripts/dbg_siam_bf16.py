import sys; sys.path.insert(0, '.')
import numpy as np, torch
from kurosiwo_b200.siam_unet import SiamUnet_conc, SiamUnet_diff
from kurosiwo_b200.lib import IMPL_SIMT
from oracle import siam_oracle, weights
DEV = "cuda:0"
for kind, H in (("conc", 64), ("conc", 128)):
    N, W, seed = 2, H, 33
    sd_np = siam_oracle.make_state(seed, 2, 3, kind)
    x1, x2, mask = (torch.from_numpy(a) for a in weights.make_batch(seed, N, H, W))
    sd = siam_oracle.to_torch_state(sd_np)
    loss_o, out_o, grads_o = siam_oracle.train_step(sd, x1, x2, mask, kind, masks=None)
    res = {}
    for impl in ("auto", "simt"):
        m = (SiamUnet_conc if kind == "conc" else SiamUnet_diff)(2, 3, precision="bf16")
        m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
        m = m.to(DEV).train(); m.dropout_p = 0.0
        eng = m.engine(x1.to(DEV))
        if impl == "simt": eng.conv_impl = IMPL_SIMT
        eng.init_training(lr=0.0)
        eng.train_step(x1.to(DEV), x2.to(DEV), mask.to(DEV))
        res[impl] = {n: eng.params.g(n).cpu().view(eng.params.offsets[n][1]).clone() for n in eng.params.names}
    print(f"== {kind} H={H}")
    for n in res["auto"]:
        if n.endswith("bias") and n.startswith("conv") and n != "conv11d.bias": continue
        go = grads_o[n]; d = go.norm().item() + 1e-30
        print(f"{n:18s} auto-vs-oracle {((res['auto'][n]-go).norm()/d):.4f}  simt-vs-oracle {((res['simt'][n]-go).norm()/d):.4f}  auto-vs-simt {((res['auto'][n]-res['simt'][n]).norm()/d):.4f}")
