"""-m gpu: the ChangeFormer CUDA kernels (csrc/cformer.cu) against the same op contracts evaluated on the CPU (tests/shadow_ops.py:
torch conv2d / softmax / interpolate semantics) on identical seeded inputs."""
import pytest
import torch

from gpu_util import mirror, rand_view, rel_l2
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


def _r(shape, dtype, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k,s,p,cin,cout,H", [(7, 4, 3, 2, 64, 32), (7, 2, 3, 64, 128, 16), (8, 8, 0, 64, 64, 16), (2, 2, 0, 40, 24, 14), (3, 1, 1, 16, 8, 9)])
def test_strided_conv_family(ops, sh, dtype, k, s, p, cin, cout, H):
    g = torch.Generator().manual_seed(1)
    N, W = 2, H + 4
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x, _ = rand_view(N, H, W, cin, dtype, DEV, gen=g)
    y, _ = rand_view(N, Ho, Wo, cout, dtype, DEV, gen=g)
    w = _r(k * k * cout * cin, dtype, g, (cin * k * k) ** -0.5)
    b = _r(cout, torch.float32, g)
    cx, cy = mirror(x), mirror(y)
    ops.conv2d_strided(N, H, W, Ho, Wo, k, s, p, x, w.to(DEV), b.to(DEV), y)
    sh.conv2d_strided(N, H, W, Ho, Wo, k, s, p, cx, w, b, cy)
    assert rel_l2(y.base.float(), cy.base.float()) < _tol(dtype)
    dy, _ = rand_view(N, Ho, Wo, cout, dtype, DEV, gen=g)
    dx, _ = rand_view(N, H, W, cin, dtype, DEV, gen=g)
    cdy, cdx = mirror(dy), mirror(dx)
    for acc in (False, True):
        ops.conv2d_strided_dgrad(N, H, W, Ho, Wo, k, s, p, dy, w.to(DEV), dx, acc)
        sh.conv2d_strided_dgrad(N, H, W, Ho, Wo, k, s, p, cdy, w, cdx, acc)
        assert rel_l2(dx.base.float(), cdx.base.float()) < _tol(dtype)
    dw = torch.randn(k * k * cout * cin, generator=g)
    cdw = dw.clone()
    dwd = dw.to(DEV)
    for acc in (True, False):
        ops.conv2d_strided_wgrad(N, H, W, Ho, Wo, k, s, p, x, dy, dwd, acc)
        sh.conv2d_strided_wgrad(N, H, W, Ho, Wo, k, s, p, cx, cdy, cdw, acc)
        assert rel_l2(dwd, cdw) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,Nq,Nk,heads,dh", [(2, 64, 49, 1, 64), (3, 196, 49, 4, 80), (2, 49, 49, 8, 64), (1, 100, 17, 2, 32), (2, 784, 49, 2, 64), (3, 196, 49, 5, 64), (4, 3136, 49, 1, 64)])
def test_xattention_fwd_bwd(ops, sh, dtype, B, Nq, Nk, heads, dh):
    g = torch.Generator().manual_seed(2)
    inner = heads * dh
    q, kv = _r((B * Nq, inner), dtype, g), _r((B * Nk, 2 * inner), dtype, g)
    scale = dh ** -0.5
    out, probs = torch.zeros(B * Nq, inner, dtype=dtype), torch.zeros(B * heads * Nq * Nk, dtype=dtype)
    sh.xattention_fwd(B, Nq, Nk, heads, dh, q, kv, scale, out, probs)
    outd, probsd = torch.ones(B * Nq, inner, dtype=dtype, device=DEV), torch.ones(B * heads * Nq * Nk, dtype=dtype, device=DEV)
    ops.xattention_fwd(B, Nq, Nk, heads, dh, q.to(DEV), kv.to(DEV), scale, outd, probsd)
    assert rel_l2(probsd.float(), probs.float()) < _tol(dtype)
    assert rel_l2(outd.float(), out.float()) < _tol(dtype)
    if dtype == torch.bfloat16 and dh == 64:
        # the tcgen05 variant (TMA -> UMMA -> TMEM softmax -> UMMA; opt-in, csrc/xattention_tc.cu) against the same shadow and, within
        # one bf16 ulp of the probabilities, against the mma.sync kernel; with dropout both draw the same stateless masks
        try:
            ops.set_option("xatt_umma", 1)
            outu, probsu = torch.ones_like(outd), torch.ones_like(probsd)
            ops.xattention_fwd(B, Nq, Nk, heads, dh, q.to(DEV), kv.to(DEV), scale, outu, probsu)
            assert rel_l2(probsu.float(), probs.float()) < _tol(dtype) and rel_l2(outu.float(), out.float()) < _tol(dtype)
            assert (probsu.float() - probsd.float()).abs().max().item() <= 2.0 ** -8
            step = torch.zeros(1, dtype=torch.int32, device=DEV)
            od, ou = torch.zeros_like(outd), torch.zeros_like(outd)
            ops.set_option("xatt_umma", 0)
            ops.xattention_fwd(B, Nq, Nk, heads, dh, q.to(DEV), kv.to(DEV), scale, od, probsd, 0.25, 777, step, 5)
            ops.set_option("xatt_umma", 1)
            ops.xattention_fwd(B, Nq, Nk, heads, dh, q.to(DEV), kv.to(DEV), scale, ou, probsu, 0.25, 777, step, 5)
            assert rel_l2(ou.float(), od.float()) < 2e-2
        finally:
            ops.set_option("xatt_umma", 0)
    dout = _r((B * Nq, inner), dtype, g)
    dq, dkv = torch.zeros_like(q), torch.zeros(B * Nk * 2 * inner)
    sh.xattention_bwd(B, Nq, Nk, heads, dh, q, kv, probs, dout, scale, dq, dkv)
    dqd, dkvd = torch.ones_like(q, device=DEV), torch.ones(B * Nk * 2 * inner, device=DEV)
    ops.xattention_bwd(B, Nq, Nk, heads, dh, q.to(DEV), kv.to(DEV), probs.to(DEV), dout.to(DEV), scale, dqd, dkvd)
    assert rel_l2(dqd.float(), dq.float()) < _tol(dtype)
    # fp32 storage: CUDA-core kernels, fp32 operands.  bf16 storage with dh = 64: the mma.sync kernel feeds dS and the dropped P to the
    # tensor cores as bf16 (as every bf16 attention backward does), so dK / dV carry bf16 operand rounding (measured 1.1e-3)
    assert rel_l2(dkvd, dkv) < (1e-4 if dtype == torch.float32 or dh != 64 else 4e-3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("simple", [0, 1, 2, 3])
@pytest.mark.parametrize("N,H,W,C", [(2, 9, 7, 40), (2, 14, 14, 256), (1, 7, 7, 2048), (3, 1, 1, 8), (1, 2, 5, 520),
                                     (8, 56, 56, 64), (20, 30, 23, 264)])      # the last two: several tiles per CTA (double-buffered pipeline), ragged tiles
def test_dwconv_fwd_bwd(ops, sh, dtype, N, H, W, C, simple):
    # 0 = shared-memory tile kernels for bf16 (default; fp32 storage runs the 2x2-block kernels), 1 = one output per thread, 2 = 2x2 blocks,
    # 3 = the tile kernels also for the weight gradient of images under 100 pixels (default: 2x2 blocks there)
    ops.set_option("dwconv_simple", simple)      # reset by the autouse fixture
    g = torch.Generator().manual_seed(3)
    x, dy = _r((N * H * W, C), dtype, g), _r((N * H * W, C), dtype, g)
    w9, b = _r((9, C), torch.float32, g, 0.3), _r(C, torch.float32, g)
    y = torch.zeros_like(x)
    sh.dwconv3x3_fwd(N, H, W, x, w9, b, y)
    yd = torch.zeros_like(x, device=DEV)
    ops.dwconv3x3_fwd(N, H, W, x.to(DEV), w9.to(DEV), b.to(DEV), yd)
    assert rel_l2(yd.float(), y.float()) < _tol(dtype)
    dx, dw9, db = torch.zeros_like(x), torch.zeros(9, C), torch.zeros(C)
    sh.dwconv3x3_bwd(N, H, W, x, dy, w9, dx, dw9, db)
    dxd, dw9d, dbd = torch.zeros_like(x, device=DEV), torch.zeros(9, C, device=DEV), torch.zeros(C, device=DEV)
    ops.dwconv3x3_bwd(N, H, W, x.to(DEV), dy.to(DEV), w9.to(DEV), dxd, dw9d, dbd)
    assert rel_l2(dxd.float(), dx.float()) < _tol(dtype)
    assert rel_l2(dw9d, dw9) < 1e-4 and rel_l2(dbd, db) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("wide", [0, 1])
@pytest.mark.parametrize("Hi,Ho", [(7, 14), (7, 56), (14, 56), (5, 12)])
def test_bilinear_nhwc_fwd_bwd(ops, sh, dtype, Hi, Ho, wide):
    ops.set_option("cf_scalar", wide)           # 1 = 64-bit index decomposition; reset by the autouse fixture
    g = torch.Generator().manual_seed(4)
    N, C = 2, 24
    Wi, Wo = Hi + 1, (Hi + 1) * Ho // Hi
    src, dst0 = _r((N * Hi * Wi, C), dtype, g), _r((N * Ho * Wo, C), dtype, g)
    for acc in (False, True):
        dst, dstd = dst0.clone(), dst0.to(DEV)
        sh.bilinear_nhwc_fwd(N, Hi, Wi, Ho, Wo, src, dst, acc)
        ops.bilinear_nhwc_fwd(N, Hi, Wi, Ho, Wo, src.to(DEV), dstd, acc)
        assert rel_l2(dstd.float(), dst.float()) < _tol(dtype)
        ds, dsd = src.clone(), src.to(DEV)
        sh.bilinear_nhwc_bwd(N, Hi, Wi, Ho, Wo, dst0, ds, acc)
        ops.bilinear_nhwc_bwd(N, Hi, Wi, Ho, Wo, dst0.to(DEV), dsd, acc)
        assert rel_l2(dsd.float(), ds.float()) < _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_relu_and_sigmoid_head(ops, sh, dtype):
    g = torch.Generator().manual_seed(5)
    x, gg = _r((50, 64), dtype, g), _r((50, 64), dtype, g)
    y, yd = torch.zeros_like(x), torch.zeros_like(x, device=DEV)
    sh.relu_fwd(x, y); ops.relu_fwd(x.to(DEV), yd)
    assert torch.equal(yd.cpu(), y)
    dx, dxd = torch.zeros_like(x), torch.zeros_like(x, device=DEV)
    sh.relu_bwd(y, gg, dx); ops.relu_bwd(yd, gg.to(DEV), dxd)
    assert torch.equal(dxd.cpu(), dx)
    N, H, W, K = 2, 6, 10, 3
    # (C, ctot, c0): dense 16 channels and an aligned 8-channel slice (16-byte stores), a 12-channel slice at channel 4 (scalar stores);
    # wide = 1: the 64-bit index arithmetic (`cf_scalar`)
    for C, ctot, c0 in ((16, 16, 0), (8, 24, 8), (12, 16, 4)):
        for wide in (0, 1):
            ops.set_option("cf_scalar", wide)
            z, _ = rand_view(N, H, W, C, dtype, DEV, ctot=ctot, c0=c0, gen=g)
            cz = mirror(z)
            out, outd = torch.zeros(N, K, H, W), torch.zeros(N, K, H, W, device=DEV)
            sh.sigmoid_head_fwd(cz, K, out); ops.sigmoid_head_fwd(z, K, outd)
            assert rel_l2(outd, out) < 1e-6
            dout = torch.randn(N, K, H, W, generator=g)
            dz, _ = rand_view(N, H, W, C, dtype, DEV, ctot=ctot, c0=c0, gen=g)
            cdz = mirror(dz)
            sh.sigmoid_head_bwd(out, dout, K, cdz); ops.sigmoid_head_bwd(outd, dout.to(DEV), K, dz)
            assert rel_l2(dz.base.float(), cdz.base.float()) < _tol(dtype)     # the whole buffer: channels outside the slice untouched


def test_permute_table_scale(ops):
    """ResidualBlock's `* 0.1` is folded into packed weights through the permute table's value multiplier."""
    src = torch.arange(24, dtype=torch.float32, device=DEV)
    dst = torch.zeros(24, dtype=torch.float32, device=DEV)
    table = ops.make_permute_table([(src, dst, (4, 6), (1, 4), 0, None, 0, 0.1)], DEV)
    ops.permute_cast_table(table)
    assert torch.allclose(dst.view(4, 6), 0.1 * src.view(6, 4).t())


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("wide", [0, 1])
@pytest.mark.parametrize("k,s,p,cin,H", [(7, 4, 3, 2, 32), (7, 2, 3, 64, 16), (8, 8, 0, 64, 16), (2, 2, 0, 40, 14), (3, 1, 1, 16, 9)])
def test_im2col_col2im(ops, sh, dtype, k, s, p, cin, H, wide):
    """ks_im2col is pure data movement (bit-exact, also on a strided source view); ks_col2im is its adjoint in gather form.
    wide = 1: the 64-bit index decomposition (`cf_scalar`; the default takes 32-bit arithmetic whenever the item count is < 2^31)."""
    ops.set_option("cf_scalar", wide)           # reset by the autouse fixture
    g = torch.Generator().manual_seed(4)
    N, W = 2, H + 4
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    Kp = (k * k * cin + 63) // 64 * 64
    x, _ = rand_view(N, H, W, cin, dtype, DEV, ctot=cin + (8 if cin % 8 == 0 else 0), c0=8 if cin % 8 == 0 else 0, gen=g)
    cx = mirror(x)
    col_c = _r((N * Ho * Wo, Kp), dtype, g)
    col_d = col_c.to(DEV)
    ops.im2col(N, H, W, Ho, Wo, k, s, p, x, col_d, Kp)
    sh.im2col(N, H, W, Ho, Wo, k, s, p, cx, col_c, Kp)
    assert torch.equal(col_d.cpu(), col_c)                             # incl. the untouched pad columns
    if cin % 8 == 0:
        dcol = _r((N * Ho * Wo, Kp), dtype, g)
        dx, _ = rand_view(N, H, W, cin, dtype, DEV, ctot=cin + 8, c0=8, gen=g)
        cdx = mirror(dx)
        for acc in (False, True):
            ops.col2im(N, H, W, Ho, Wo, k, s, p, dcol.to(DEV), Kp, dx, acc)
            sh.col2im(N, H, W, Ho, Wo, k, s, p, dcol, Kp, cdx, acc)
            assert rel_l2(dx.base.float(), cdx.base.float()) < _tol(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_dropout_branch_vector_kernels_equal_scalar(ops, sh, dtype):
    """The 16-byte Dropout / DropPath kernels (default) draw the same per-element masks and evaluate the same expressions as the scalar
    kernels (`cf_scalar` = 1): bit-identical results; both against the numpy restatement of the RNG (tests/shadow_ops.py)."""
    g = torch.Generator().manual_seed(12)
    B, per = 5, 40 * 24                                              # per_sample % 8 == 0: a 16-byte vector never straddles two samples
    n = B * per
    step = torch.tensor([7], dtype=torch.int32)
    x, t = _r(n, dtype, g), _r(n, dtype, g)
    dp = torch.tensor([1.25, 0.0, 1.25, 1.25, 0.0])
    res = {}
    for scalar in (0, 1):
        ops.set_option("cf_scalar", scalar)
        out = []
        for p, dpv in ((0.1, dp), (0.1, None), (0.0, dp)):
            y = torch.zeros(n, dtype=dtype, device=DEV)
            ops.dropout_apply(x.to(DEV), y, max(p, 0.05), 99, step.to(DEV), 5)
            xa = x.to(DEV)
            ops.branch_add(xa, t.to(DEV), per, p, None if dpv is None else dpv.to(DEV), 99, step.to(DEV), 6)
            dt = torch.zeros(n, dtype=dtype, device=DEV)
            ops.branch_scale(x.to(DEV), dt, per, p, None if dpv is None else dpv.to(DEV), 99, step.to(DEV), 6)
            xin = x.to(DEV)
            ops.dropout_apply(xin, xin, 0.1, 99, step.to(DEV), 7)     # in place, as the engine calls it
            out += [y.cpu(), xa.cpu(), dt.cpu(), xin.cpu()]
        res[scalar] = out
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
    ys = torch.zeros_like(x)
    sh.dropout_apply(x, ys, 0.1, 99, step, 5)
    assert torch.equal(res[0][0], ys)
    xs = x.clone()
    sh.branch_add(xs, t, per, 0.1, dp, 99, step, 6)
    assert torch.allclose(res[0][1].float(), xs.float(), rtol=1e-2 if dtype == torch.bfloat16 else 1e-6, atol=1e-2 if dtype == torch.bfloat16 else 1e-6)
