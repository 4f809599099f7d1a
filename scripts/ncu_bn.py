"""ncu target: a few launches of the BatchNorm passes at the SNUNet level-0 shape."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kurosiwo_b200.lib import CudaOps, View
ops = CudaOps(); dev = "cuda:0"; bf = torch.bfloat16; N, H, C = 64, 224, 32
mk = lambda: View.alloc(N, H, H, C, bf, dev, zero=False)
y, out, dout, dy = mk(), mk(), mk(), mk()
for v in (y, dout): v.base.normal_()
sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev) * 0.1; mu = torch.zeros(C, device=dev); rs = torch.ones(C, device=dev)
gamma = torch.ones(C, device=dev); sums = torch.zeros(2 * C, dtype=torch.float64, device=dev); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
for _ in range(3):
    ops.bn_act(y, sc, sh, None, True, out, None)
    ops.bn_bwd_reduce(dout, out, y, None, None, mu, rs, sums)
    ops.bn_bwd_apply(dout, True, y, None, None, mu, rs, gamma, sums, float(N * H * H), None, dy, dg, db, None, False)
    out.base.copy_(y.base)
torch.cuda.synchronize()
