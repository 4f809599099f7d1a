"""HBM roofline of the fused CE+Dice kernel at the BASELINE batch (bs=64, 224x224, 3 classes): rotating buffer sets defeat the L2."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import CudaOps
ops = CudaOps(); dev = "cuda:0"; N, H, W = 64, 224, 224
sets = []
for i in range(6):
    g = torch.Generator(device=dev).manual_seed(i)
    sets.append((torch.randn(N, 3, H, W, device=dev, generator=g), torch.randint(0, 4, (N, H, W), device=dev, generator=g),
                 torch.empty(N, 3, H, W, device=dev), torch.empty(N, H, W, dtype=torch.uint8, device=dev)))
w = torch.ones(3, device=dev); loss3 = torch.zeros(3, device=dev); ws = ops.ce_dice_workspace(N, dev)
def run(i):
    z, y, dz, pr = sets[i % len(sets)]
    ops.ce_dice(z, y, w, 3, 1.0, loss3, dz, pr, ws)
for i in range(6): run(i)
torch.cuda.synchronize()
reps = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps): run(i)
e1.record(); torch.cuda.synchronize()
us_eager = e0.elapsed_time(e1) / reps * 1e3
# the same 30 calls as ONE CUDA graph: the device-side time per call, free of the host's launch rate (3 launches per call)
g = torch.cuda.CUDAGraph()
s_ = torch.cuda.Stream()
with torch.cuda.stream(s_):
    for i in range(3): run(i)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s_):
        for i in range(reps): run(i)
    g.replay(); torch.cuda.synchronize()
    e0.record(s_)
    for _ in range(5): g.replay()
    e1.record(s_); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / (5 * reps) * 1e3
alg = N * H * W * 32 + N * H * W          # 12 B logits + 8 B label + 12 B gradient per pixel, + 1 B argmax
print(f"ce_dice bs={N}: {us:.1f} us per call as a CUDA graph ({us_eager:.1f} us eager, host-launch bound; no memset: self-cleaning workspace), "
      f"algorithmic {alg/1e6:.1f} MB -> {alg/us/1e6:.2f} TB/s = {alg/us/1e6/6.4549:.2f} of the measured 6454.9 GB/s; loss {loss3.tolist()}")
