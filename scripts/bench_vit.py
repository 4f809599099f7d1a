"""Timing probe of the FloodViT (ViT-B/16, 6 channels, linear head) training step at bs=64 on one B200 (BASELINE.json configs[3] shape)."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.vision_transformer import FinetunerSegmentation, ViT
from kurosiwo_b200 import synthetic
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda:0"
torch.manual_seed(0)
enc = ViT(image_size=224, patch_size=16, num_classes=3, dim=768, depth=12, heads=12, mlp_dim=3072, channels=6)
m = FinetunerSegmentation(encoder=enc, configs={"mlp": False, "decoder": False, "num_classes": 3, "finetuning_patch_size": 16}).to(dev).train()
b = synthetic.make_batch(999, bs)
img = torch.cat((b[2], b[6], b[9]), 1).to(dev); mask = b[3].to(dev)
eng = m.engine(img); eng.init_training(lr=1e-4)
for _ in range(2): eng.train_step(img, mask)
torch.cuda.synchronize()
l0 = eng.ops.launches; eng.train_step(img, mask); calls = eng.ops.launches - l0
# per-op timing of one eager step
import collections
times = collections.OrderedDict()
ops = eng.ops
orig = {}
def wrap(name):
    f = getattr(ops, name)
    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **k); e1.record()
        times.setdefault(name, []).append((e0, e1)); return r
    orig[name] = f; setattr(ops, name, g)
for n in ("conv2d", "conv2d_wgrad", "attention_fwd", "attention_bwd", "layernorm_fwd", "layernorm_bwd", "gelu_fwd", "gelu_bwd", "channel_sum",
          "permute_cast_table", "patchify_ln", "patchify_ln_bwd", "bilinear_up_fwd", "bilinear_up_bwd", "ce_dice", "adam_step", "vit_assemble", "vit_assemble_bwd", "zero_"):
    wrap(n)
eng.train_step(img, mask); torch.cuda.synchronize()
for n, f in orig.items(): setattr(ops, n, f)
tot = 0
for n, ev in times.items():
    ms = sum(a.elapsed_time(b_) for a, b_ in ev); tot += ms
    print(f"{n:20s} n={len(ev):4d} {ms:8.3f} ms")
print("sum of op times", tot)
step = eng.capture(img, mask)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gf = 3 * (33.93 + 1.43)
print(f"FloodViT-B bs={bs}: {ms:.2f} ms/step, {bs / ms * 1e3:.0f} patches/s, {bs * gf / ms:.0f} TFLOP/s-equivalent (106 GF/patch), {calls} C-ABI calls/step, loss {eng.loss3[0].item():.4f}")
