"""Depth-wise 3x3 conv kernels (ChangeFormer Mlp, changeformer.py:126-133) at the four encoder stage shapes of the bs=32 step (2 x 32 images):
time per call and algorithmic GB/s (x read once + y written once; backward: x, dy read + dx written) for the three kernel families
(`dwconv_simple` = 0 shared-memory tiles / 1 one output per thread / 2 register 2x2 blocks), and the bit-identity of the forward and
data-gradient results across them.   python scripts/bench_dwconv.py [N]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import CudaOps

ops = CudaOps(); dev = "cuda:0"; bf = torch.bfloat16
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
SHAPES = [(56, 256), (28, 512), (14, 1280), (7, 2048)]
NAMES = {0: "tile", 1: "simple", 2: "block"}
NBUF = 4                                            # rotating operand sets: 4 x (3 x 103 MB) at stage 1 - nothing is served from the 126 MB L2


def timed(fn, reps):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


tot = {v: [0.0, 0.0] for v in NAMES}
for H, C in SHAPES:
    g = torch.Generator(device=dev).manual_seed(H)
    xs = [torch.randn(N * H * H, C, device=dev, generator=g).to(bf) for _ in range(NBUF)]
    dys = [torch.randn(N * H * H, C, device=dev, generator=g).to(bf) for _ in range(NBUF)]
    outs = [torch.empty_like(xs[0]) for _ in range(NBUF)]
    w9 = torch.randn(9, C, device=dev, generator=g) * 0.3; b = torch.randn(C, device=dev, generator=g)
    dw9, db = torch.zeros(9, C, device=dev), torch.zeros(C, device=dev)
    nbytes = xs[0].numel() * 2
    ref = {}
    for v in NAMES:
        ops.set_option("dwconv_simple", v)
        y = torch.empty_like(xs[0]); dx = torch.empty_like(xs[0]); dw9.zero_(); db.zero_()
        ops.dwconv3x3_fwd(N, H, H, xs[0], w9, b, y)
        ops.dwconv3x3_bwd(N, H, H, xs[0], dys[0], w9, dx, dw9, db)
        torch.cuda.synchronize()
        if v == 0:
            ref = dict(y=y, dx=dx, dw9=dw9.clone(), db=db.clone())
        else:
            same = torch.equal(y, ref["y"]) and torch.equal(dx, ref["dx"])
            rw = float((dw9 - ref["dw9"]).norm() / ref["dw9"].norm()); rb = float((db - ref["db"]).norm() / ref["db"].norm())
            print(f"  {H}x{H} C={C}: tile vs {NAMES[v]}: y/dx bit-identical {same}, dw9 rel {rw:.2e}, dbias rel {rb:.2e}")
        tf = timed(lambda i: ops.dwconv3x3_fwd(N, H, H, xs[i % NBUF], w9, b, outs[i % NBUF]), 20)
        tb = timed(lambda i: ops.dwconv3x3_bwd(N, H, H, xs[i % NBUF], dys[i % NBUF], w9, outs[i % NBUF], dw9, db), 20)
        tot[v][0] += tf; tot[v][1] += tb
        print(f"{H}x{H} C={C} N={N} {NAMES[v]:>6}: fwd {tf:7.1f} us = {2 * nbytes / tf / 1e3:6.0f} GB/s   bwd (dgrad + wgrad) {tb:7.1f} us = {5 * nbytes / tb / 1e3:6.0f} GB/s")
for v in NAMES:
    print(f"sum over the four stage shapes, {NAMES[v]:>6}: fwd {tot[v][0]:.1f} us, bwd {tot[v][1]:.1f} us")
ops.reset_options()
