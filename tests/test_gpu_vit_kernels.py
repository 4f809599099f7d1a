"""-m gpu: the ViT / FloodViT-head CUDA kernels (csrc/vit.cu) against the same op contracts evaluated on the CPU
(tests/shadow_ops.py: torch layer_norm / softmax / gelu / interpolate semantics) on identical seeded inputs."""
import pytest
import torch

from gpu_util import rel_l2
from shadow_ops import ShadowOps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DTYPES = [torch.float32, torch.bfloat16]


@pytest.fixture(scope="module")
def ops():
    from kurosiwo_b200.lib import CudaOps
    return CudaOps()


@pytest.fixture(scope="module")
def sh():
    return ShadowOps()


def _tol(dtype):
    return 3e-5 if dtype == torch.float32 else 8e-3


def _r(shape, dtype, g, scale=1.0, shift=0.0):
    return (torch.randn(shape, generator=g) * scale + shift).to(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,C", [(37, 64), (1003, 64), (130, 128), (75, 200), (64, 256), (77, 320), (416, 768), (50, 1536), (9, 2048), (5, 8)])
def test_layernorm_fwd(ops, sh, dtype, rows, C):
    g = torch.Generator().manual_seed(1)
    x = _r((rows, C), dtype, g, 2.0, 0.5)
    gamma, beta = _r(C, torch.float32, g, 0.2, 1.0), _r(C, torch.float32, g, 0.2)
    y, cp = torch.zeros_like(x), torch.zeros_like(x)
    mu, rs = torch.zeros(rows), torch.zeros(rows)
    sh.layernorm_fwd(x, gamma, beta, 1e-5, y, mu, rs, cp)
    xd = x.to(DEV)
    yd, cpd, mud, rsd = torch.zeros_like(xd), torch.zeros_like(xd), torch.zeros(rows, device=DEV), torch.zeros(rows, device=DEV)
    ops.layernorm_fwd(xd, gamma.to(DEV), beta.to(DEV), 1e-5, yd, mud, rsd, cpd)
    assert rel_l2(yd.float(), y.float()) < _tol(dtype)
    assert rel_l2(mud, mu) < 1e-5 and rel_l2(rsd, rs) < 1e-5
    assert torch.equal(cpd.cpu(), x)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,C,acc", [(37, 64, False), (1001, 64, True), (130, 128, False), (77, 320, True), (300, 512, False), (416, 768, True), (100, 1024, False), (5, 8, False)])
def test_layernorm_bwd(ops, sh, dtype, rows, C, acc):
    g = torch.Generator().manual_seed(2)
    x, dy = _r((rows, C), dtype, g, 2.0, 0.5), _r((rows, C), dtype, g)
    gamma = _r(C, torch.float32, g, 0.2, 1.0)
    xf = x.float()
    mu = xf.mean(1)
    rs = torch.rsqrt(((xf - mu[:, None]) ** 2).mean(1) + 1e-5)
    dx0 = _r((rows, C), dtype, g)
    dx, dg, db = dx0.clone(), torch.zeros(C), torch.zeros(C)
    sh.layernorm_bwd(dy, x, mu, rs, gamma, dx, acc, dg, db)
    dxd, dgd, dbd = dx0.to(DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.layernorm_bwd(dy.to(DEV), x.to(DEV), mu.to(DEV), rs.to(DEV), gamma.to(DEV), dxd, acc, dgd, dbd)
    assert rel_l2(dxd.float(), dx.float()) < _tol(dtype)
    assert rel_l2(dgd, dg) < 1e-4 and rel_l2(dbd, db) < 1e-4
    # parameter-gradient-only mode
    dgd.zero_(); dbd.zero_()
    ops.layernorm_bwd(dy.to(DEV), x.to(DEV), mu.to(DEV), rs.to(DEV), gamma.to(DEV), None, False, dgd, dbd)
    assert rel_l2(dgd, dg) < 1e-4 and rel_l2(dbd, db) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,Cc,HW", [(2, 6, 64), (3, 6, 224), (2, 3, 32)])
def test_patchify_ln_fwd_bwd(ops, sh, dtype, B, Cc, HW):
    g = torch.Generator().manual_seed(3)
    n = (HW // 16) ** 2
    Tp = (n + 1 + 15) // 16 * 16
    PD = 256 * Cc
    img = _r((B, Cc, HW, HW), torch.float32, g, 1.5, 0.3)
    gamma, beta = _r(PD, torch.float32, g, 0.2, 1.0), _r(PD, torch.float32, g, 0.2)
    out, mu, rs = torch.zeros(B * Tp, PD, dtype=dtype), torch.zeros(B * Tp), torch.zeros(B * Tp)
    sh.patchify_ln(img, Tp, gamma, beta, 1e-5, out, mu, rs)
    outd, mud, rsd = torch.zeros(B * Tp, PD, dtype=dtype, device=DEV), torch.zeros(B * Tp, device=DEV), torch.zeros(B * Tp, device=DEV)
    ops.patchify_ln(img.to(DEV), Tp, gamma.to(DEV), beta.to(DEV), 1e-5, outd, mud, rsd)
    assert rel_l2(outd.float(), out.float()) < _tol(dtype)
    assert rel_l2(mud, mu) < 1e-5 and rel_l2(rsd, rs) < 1e-5
    dy = _r((B * Tp, PD), dtype, g)
    dg, db = torch.zeros(PD), torch.zeros(PD)
    sh.patchify_ln_bwd(img, Tp, mu, rs, dy, dg, db)
    dgd, dbd = torch.zeros(PD, device=DEV), torch.zeros(PD, device=DEV)
    ops.patchify_ln_bwd(img.to(DEV), Tp, mud, rsd, dy.to(DEV), dgd, dbd)
    assert rel_l2(dgd, dg) < 1e-4 and rel_l2(dbd, db) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_vit_assemble_fwd_bwd(ops, sh, dtype):
    g = torch.Generator().manual_seed(4)
    B, T, Tp, D = 3, 17, 32, 64
    e = _r((B * Tp, D), dtype, g)
    cls, pos = _r(D, torch.float32, g), _r((T, D), torch.float32, g)
    x0 = torch.zeros(B * Tp, D, dtype=dtype)
    sh.vit_assemble(B, T, Tp, e, cls, pos, x0)
    x0d = torch.ones(B * Tp, D, dtype=dtype, device=DEV)
    ops.vit_assemble(B, T, Tp, e.to(DEV), cls.to(DEV), pos.to(DEV), x0d)
    assert rel_l2(x0d.float(), x0.float()) < _tol(dtype)
    dx0 = _r((B * Tp, D), dtype, g)
    de, dcls, dpos = torch.ones(B * Tp, D, dtype=dtype), torch.zeros(D), torch.zeros(T, D)
    sh.vit_assemble_bwd(B, T, Tp, dx0, de, dcls, dpos)
    ded, dclsd, dposd = torch.ones(B * Tp, D, dtype=dtype, device=DEV), torch.zeros(D, device=DEV), torch.zeros(T, D, device=DEV)
    ops.vit_assemble_bwd(B, T, Tp, dx0.to(DEV), ded, dclsd, dposd)
    assert torch.equal(ded.cpu(), de)
    assert rel_l2(dclsd, dcls) < 1e-5 and rel_l2(dposd, dpos) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,T,Tp,heads", [(2, 17, 32, 2), (3, 197, 208, 4), (1, 197, 208, 12)])
def test_attention_fwd_bwd(ops, sh, dtype, B, T, Tp, heads):
    g = torch.Generator().manual_seed(5)
    dh, inner = 64, heads * 64
    qkv = _r((B * Tp, 3 * inner), dtype, g)
    scale = dh ** -0.5
    out, probs = torch.zeros(B * Tp, inner, dtype=dtype), torch.zeros(B * heads * Tp * Tp, dtype=dtype)
    sh.attention_fwd(B, T, Tp, heads, dh, qkv, scale, out, probs)
    qd = qkv.to(DEV)
    outd, probsd = torch.ones(B * Tp, inner, dtype=dtype, device=DEV), torch.ones(B * heads * Tp * Tp, dtype=dtype, device=DEV)
    ops.attention_fwd(B, T, Tp, heads, dh, qd, scale, outd, probsd)
    assert rel_l2(probsd.float(), probs.float()) < _tol(dtype)
    assert rel_l2(outd.float(), out.float()) < _tol(dtype)
    dout = _r((B * Tp, inner), dtype, g)
    dqkv = torch.zeros_like(qkv)
    sh.attention_bwd(B, T, Tp, heads, dh, qkv, probs, dout, scale, dqkv, None)
    dqkvd, dsd = torch.ones_like(qd), torch.zeros(B * heads * Tp * Tp, dtype=dtype, device=DEV)
    ops.attention_bwd(B, T, Tp, heads, dh, qd, probs.to(DEV), dout.to(DEV), scale, dqkvd, dsd)
    assert rel_l2(dqkvd.float(), dqkv.float()) < (_tol(dtype) if dtype == torch.float32 else 2e-2)
    pad = dqkvd.view(B, Tp, -1)[:, T:]
    assert float(pad.float().abs().max()) == 0.0 if pad.numel() else True


@pytest.mark.parametrize("dtype", DTYPES)
def test_gelu_fwd_bwd(ops, sh, dtype):
    g = torch.Generator().manual_seed(6)
    u, dh = _r((40, 136), dtype, g, 2.0), _r((40, 136), dtype, g)
    h, du = torch.zeros_like(u), torch.zeros_like(u)
    sh.gelu_fwd(u, h); sh.gelu_bwd(u, dh, du)
    hd, dud = torch.zeros_like(u, device=DEV), torch.zeros_like(u, device=DEV)
    ops.gelu_fwd(u.to(DEV), hd); ops.gelu_bwd(u.to(DEV), dh.to(DEV), dud)
    assert rel_l2(hd.float(), h.float()) < _tol(dtype) and rel_l2(dud.float(), du.float()) < _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("G,Ho", [(14, 224), (4, 64), (2, 32)])
def test_bilinear_up_fwd_bwd(ops, sh, dtype, G, Ho):
    g = torch.Generator().manual_seed(7)
    B, K, Cs = 2, 3, 16
    Tp = (G * G + 1 + 15) // 16 * 16
    src = _r((B * Tp, Cs), dtype, g)
    dst = torch.zeros(B, K, Ho, Ho)
    sh.bilinear_up_fwd(B, G, Tp, 1, K, Ho, Ho, src, dst)
    dstd = torch.zeros(B, K, Ho, Ho, device=DEV)
    ops.bilinear_up_fwd(B, G, Tp, 1, K, Ho, Ho, src.to(DEV), dstd)
    assert rel_l2(dstd, dst) < 1e-5
    dd = torch.randn(B, K, Ho, Ho, generator=g)
    ds = torch.ones(B * Tp, Cs, dtype=dtype)
    sh.bilinear_up_bwd(B, G, Tp, 1, K, Ho, Ho, dd, ds)
    dsd = torch.ones(B * Tp, Cs, dtype=dtype, device=DEV)
    ops.bilinear_up_bwd(B, G, Tp, 1, K, Ho, Ho, dd.to(DEV), dsd)
    assert rel_l2(dsd.float(), ds.float()) < _tol(dtype)


@pytest.mark.parametrize("M,K,N", [(416, 768, 2304), (832, 3072, 768), (208, 1536, 64)])
def test_linear_as_conv1x1_tc(ops, sh, M, K, N):
    """The token GEMMs run on the tcgen05 conv engine as 1x1 convolutions over [1, M/16, 16, C] views: fwd, dgrad, wgrad."""
    from kurosiwo_b200.lib import IMPL_TC, View
    g = torch.Generator().manual_seed(8)
    bf = torch.bfloat16
    a = _r((M, K), bf, g).to(DEV)
    w = _r((N, K), bf, g, K ** -0.5).to(DEV)
    bias = _r(N, torch.float32, g).to(DEV)
    y = torch.zeros(M, N, dtype=bf, device=DEV)
    va = View(a.view(-1), 0, 1, M // 16, 16, K, M * K, 16 * K, K)
    vy = View(y.view(-1), 0, 1, M // 16, 16, N, M * N, 16 * N, N)
    ops.conv2d(1, M // 16, 16, 1, [va], w.view(-1), bias, [vy], None, None, IMPL_TC)
    ref = a.float() @ w.float().t() + bias
    assert rel_l2(y.float(), ref) < 5e-3
    # accumulate into the destination (residual stream)
    y0 = y.clone()
    ops.conv2d(1, M // 16, 16, 1, [va], w.view(-1), bias, [vy], [True], None, IMPL_TC)
    assert rel_l2(y.float(), y0.float() + ref) < 6e-3
    # weight gradient: dw[n][k] = sum_m dy[m][n] a[m][k]
    dy = _r((M, N), bf, g).to(DEV)
    vdy = View(dy.view(-1), 0, 1, M // 16, 16, N, M * N, 16 * N, N)
    dw = torch.zeros(N * K, device=DEV)
    ops.conv2d_wgrad(1, M // 16, 16, 1, [va], [vdy], dw, False, IMPL_TC)
    assert rel_l2(dw.view(N, K), dy.float().t() @ a.float()) < 5e-3
