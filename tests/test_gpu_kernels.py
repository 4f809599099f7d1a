"""-m gpu: every CUDA kernel behind the C ABI against the same op contract evaluated on the CPU
(tests/shadow_ops.py, oracle/loss_oracle.py) on identical seeded inputs."""
import numpy as np
import pytest
import torch

from gpu_util import mirror, max_rel, rand_view, rel_l2
from shadow_ops import ShadowOps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from kurosiwo_b200.lib import CudaOps
    return CudaOps()


@pytest.fixture(scope="module")
def sh():
    return ShadowOps()


def _tol(dtype):
    return 2e-5 if dtype == torch.float32 else 6e-3


def _gen(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ks,srcspec,cout", [(3, [(2, 2, 0)], 32), (3, [(24, 40, 8), (16, 16, 0)], 48), (1, [(32, 32, 0)], 64),
                                               (3, [(96, 192, 0), (64, 64, 0)], 32)])
def test_conv_simt(ops, sh, dtype, ks, srcspec, cout):
    from kurosiwo_b200.lib import IMPL_SIMT
    g = _gen(1)
    N, H, W = 2, 12, 10
    srcs = [rand_view(N, H, W, c, dtype, DEV, ct, c0, gen=g)[0] for (c, ct, c0) in srcspec]
    cin = sum(c for c, _, _ in srcspec)
    d1, f1 = rand_view(N, H, W, cout - 16, dtype, DEV, cout + 8, 8, gen=g)
    d2, f2 = rand_view(N, H, W, 16, dtype, DEV, gen=g)
    w = (torch.randn(ks * ks * cout * cin, generator=g) * (1.0 / (cin * ks * ks)) ** 0.5).to(dtype).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    csrcs, cd1, cd2 = [mirror(s) for s in srcs], mirror(d1), mirror(d2)
    cd1.base, cd2.base = f1.base.cpu().clone(), f2.base.cpu().clone()
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    ops.conv2d(N, H, W, ks, srcs, w, b, [d1, d2], [False, True], None, IMPL_SIMT)
    sh.conv2d(N, H, W, ks, csrcs, w.cpu(), b.cpu(), [cd1, cd2], [False, True], None)
    assert rel_l2(f1.base.float(), cd1.base.float()) < _tol(dtype)
    assert rel_l2(f2.base.float(), cd2.base.float()) < _tol(dtype)
    # fused statistics (single assign destination)
    d3, f3 = rand_view(N, H, W, cout, dtype, DEV, gen=g)
    cd3 = mirror(d3)
    cstats = torch.zeros(2 * cout, dtype=torch.float64)
    ops.conv2d(N, H, W, ks, srcs, w, b, [d3], [False], stats, IMPL_SIMT)
    sh.conv2d(N, H, W, ks, csrcs, w.cpu(), b.cpu(), [cd3], [False], cstats)
    assert rel_l2(stats, cstats) < 1e-2 if dtype == torch.bfloat16 else rel_l2(stats, cstats) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ks", [3, 1])
def test_wgrad_simt(ops, sh, dtype, ks):
    from kurosiwo_b200.lib import IMPL_SIMT
    g = _gen(2)
    N, H, W = 2, 10, 12
    xs = [rand_view(N, H, W, 24, dtype, DEV, 40, 8, gen=g)[0], rand_view(N, H, W, 70, dtype, DEV, gen=g)[0]]
    dys = [rand_view(N, H, W, 20, dtype, DEV, gen=g)[0], rand_view(N, H, W, 8, dtype, DEV, 16, 8, gen=g)[0]]
    n = ks * ks * 28 * 94
    dw = torch.randn(n, generator=g).to(DEV)
    cdw = dw.cpu().clone()
    ops.conv2d_wgrad(N, H, W, ks, xs, dys, dw, True, IMPL_SIMT)
    sh.conv2d_wgrad(N, H, W, ks, [mirror(v) for v in xs], [mirror(v) for v in dys], cdw, True)
    assert rel_l2(dw, cdw) < 1e-4
    ops.conv2d_wgrad(N, H, W, ks, xs, dys, dw, False, IMPL_SIMT)
    sh.conv2d_wgrad(N, H, W, ks, [mirror(v) for v in xs], [mirror(v) for v in dys], cdw, False)
    assert rel_l2(dw, cdw) < 1e-4


@pytest.mark.parametrize("cins,couts,phases,mod", [([64], [64], True, 64), ([128], [128], True, 128), ([32], [32], False, 0),
                                                  ([64, 32], [64, 32], False, 0), ([64], [256], False, 0), ([320], [64], False, 0)])
def test_wgrad_with_fused_bias_gradient(ops, sh, cins, couts, phases, mod):
    """ks_conv2d_wgrad_bias (1x1, bf16, tcgen05): the per-channel sum of dy rides along in the weight-gradient UMMAs when the last M tile of
    Cin has a free group slot (Cin = 64, 32, 96, 320), else falls back to ks_channel_sum (Cin = 128); four ConvTranspose phase views fold
    onto one bias (mod = C).  Against the shadow, incl. accumulation over two calls."""
    from kurosiwo_b200.lib import IMPL_TC
    g = _gen(77)
    N, H, W = 3, 24, 32
    dt = torch.bfloat16
    xs = [rand_view(N, H, W, c, dt, DEV, gen=g)[0] for c in cins]
    if phases:
        _, full = rand_view(N, 2 * H, 2 * W, couts[0], dt, DEV, gen=g)
        dys = [full.phase(k // 2, k % 2) for k in range(4)]
        cfull = mirror(full)
        cdys = [cfull.phase(k // 2, k % 2) for k in range(4)]
    else:
        dys = [rand_view(N, H, W, c, dt, DEV, gen=g)[0] for c in couts]
        cdys = [mirror(v) for v in dys]
    cxs = [mirror(v) for v in xs]
    cin, cout = sum(cins), sum(v.C for v in dys)
    blen = mod if mod else cout
    dw, cdw = torch.full((cout * cin,), 3.0, device=DEV), torch.full((cout * cin,), 3.0)
    db, cdb = torch.full((blen,), 5.0, device=DEV), torch.full((blen,), 5.0)
    for acc in (False, True):
        ops.conv2d_wgrad_bias(N, H, W, 1, xs, dys, dw, db, mod, acc, acc, IMPL_TC)
        sh.conv2d_wgrad_bias(N, H, W, 1, cxs, cdys, cdw, cdb, mod, acc, acc)
        assert rel_l2(dw, cdw) < 1e-4
        assert rel_l2(db, cdb) < 1e-4, (db[:8], cdb[:8])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cin,H,W", [(2, 24, 40), (3, 17, 33), (2, 50, 70)])
def test_stem_kernels(ops, sh, dtype, cin, H, W):
    g = _gen(8)
    N = 3
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(32 * cin * 9, generator=g) * 0.3
    b = torch.randn(32, generator=g) * 0.1
    dst, fd = rand_view(N, H, W, 32, dtype, DEV, 96, 32, gen=g)
    cdst = mirror(dst)
    stats, cstats = torch.zeros(64, dtype=torch.float64, device=DEV), torch.zeros(64, dtype=torch.float64)
    ops.stem_conv3x3(x.to(DEV), w.to(DEV), b.to(DEV), dst, stats)
    sh.stem_conv3x3(x, w, b, cdst, cstats)
    assert rel_l2(fd.base.float(), cdst.base.float()) < _tol(dtype)
    assert rel_l2(stats, cstats) < (1e-5 if dtype == torch.float32 else 1e-2)
    dy, _ = rand_view(N, H, W, 32, dtype, DEV, gen=g)
    dw, cdw = torch.ones(32 * cin * 9, device=DEV), torch.ones(32 * cin * 9)
    for acc in (True, False):
        ops.stem_wgrad3x3(x.to(DEV), dy, dw, acc)
        sh.stem_wgrad3x3(x, mirror(dy), cdw, acc)
        assert rel_l2(dw, cdw) < 1e-4


def test_permute_cast(ops, sh):
    g = _gen(3)
    w = torch.randn(6 * 5 * 9, generator=g)
    for dt in (torch.float32, torch.bfloat16):
        dst = torch.zeros(9 * 6 * 5, dtype=dt, device=DEV)
        cdst = torch.zeros(9 * 6 * 5, dtype=dt)
        ops.permute_cast(w.to(DEV), dst, (9, 5, 6), (-1, 9, 45), src_offset=8)
        sh.permute_cast(w, cdst, (9, 5, 6), (-1, 9, 45), src_offset=8)
        assert torch.equal(dst.cpu(), cdst)
    x = torch.randn(2, 3, 4, 6, generator=g)
    dst = torch.zeros(2 * 4 * 6 * 3, device=DEV)
    ops.permute_cast(x.to(DEV), dst, (2, 4, 6, 3), (72, 6, 1, 24))
    assert torch.equal(dst.cpu().view(2, 4, 6, 3), x.permute(0, 2, 3, 1).contiguous())
    ops.permute_cast(x.to(DEV), dst, (2, 4, 6, 3), (72, 6, 1, 24), accumulate=True)
    assert torch.allclose(dst.cpu().view(2, 4, 6, 3), 2 * x.permute(0, 2, 3, 1))


def test_permute_cast_batched(ops, sh):
    g = _gen(9)
    srcs = [torch.randn(n, generator=g).to(DEV) for n in (9 * 40 * 24, 4 * 16 * 16, 5000, 200 * 136, 72 * 100, 8200)]
    jobs, cjobs = [], []
    specs = [((9, 24, 40), (-1, 9, 216), 8, torch.bfloat16), ((16, 4, 16), (64, 1, 4), 0, torch.float32), ((5000,), (1,), 0, torch.bfloat16),
             ((136, 200), (1, 136), 0, torch.bfloat16),        # [out=200][in=136] -> [in][out]: the tiled-transpose path (ragged tiles)
             ((72, 100), (1, 72), 0, torch.float32),           # the same, fp32 destination
             ((8192,), (1,), 8, torch.bfloat16)]               # vectorised contiguous cast (32-byte aligned source offset)
    for src, (dims, strides, off, dt) in zip(srcs, specs):
        n = 1
        for d in dims:
            n *= d
        dst, cdst = torch.zeros(n, dtype=dt, device=DEV), torch.zeros(n, dtype=dt)
        jobs.append((src, dst, dims, strides, off))
        cjobs.append((src.cpu(), cdst, dims, strides, off))
    ops.permute_cast_table(ops.make_permute_table(jobs, DEV))
    sh.permute_cast_table(sh.make_permute_table(cjobs, "cpu"))
    for (_, d, *_r), (_, cd, *_r2) in zip(jobs, cjobs):
        assert torch.equal(d.cpu(), cd)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,ctot,c0", [(32, 32, 0), (64, 192, 64), (3, 3, 0), (512, 512, 0)])
def test_bn_forward_backward(ops, sh, dtype, C, ctot, c0):
    g = _gen(4)
    N, H, W = 3, 8, 12
    y, _ = rand_view(N, H, W, C, dtype, DEV, gen=g)
    res, _ = rand_view(N, H, W, C, dtype, DEV, gen=g)
    out, fout = rand_view(N, H, W, C, dtype, DEV, ctot, c0, gen=g)
    pool, fpool = rand_view(N, H // 2, W // 2, C, dtype, DEV, gen=g)
    gamma, beta = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=DEV)
    small = [torch.zeros(C, device=DEV) for _ in range(4)]
    cy, cres, cout, cpool = mirror(y), mirror(res), mirror(out), mirror(pool)
    cout.base, cpool.base = fout.base.cpu().clone(), fpool.base.cpu().clone()
    csums, crm, crv = sums.cpu().clone(), rm.cpu().clone(), rv.cpu().clone()
    csmall = [t.cpu().clone() for t in small]
    count = float(N * H * W)
    ops.bn_stats(y, sums)
    sh.bn_stats(cy, csums)
    assert rel_l2(sums, csums) < 1e-6
    ops.bn_finalize(C, count, sums, gamma, beta, 1e-5, 0.1, rm, rv, *small)
    sh.bn_finalize(C, count, csums, gamma.cpu(), beta.cpu(), 1e-5, 0.1, crm, crv, *csmall)
    for a, b in zip(small + [rm, rv], csmall + [crm, crv]):
        assert max_rel(a, b) < 1e-5
    ops.bn_act(y, small[0], small[1], res, True, out, pool)
    sh.bn_act(cy, csmall[0], csmall[1], cres, True, cout, cpool)
    assert rel_l2(fout.base.float(), cout.base.float()) < _tol(dtype)
    assert rel_l2(fpool.base.float(), cpool.base.float()) < _tol(dtype)
    # backward, pass 1 with the mask taken from `out` (masked gradient written back in place)
    dout, fdo = rand_view(N, H, W, C, dtype, DEV, ctot, c0, gen=g)
    add_d, _ = rand_view(N, H, W, C, dtype, DEV, gen=g)
    dy, fdy = rand_view(N, H, W, C, dtype, DEV, gen=g)
    bs = torch.zeros(2 * C, dtype=torch.float64, device=DEV)
    cbs = bs.cpu().clone()
    cdout, cadd, cdy = mirror(dout), mirror(add_d), mirror(dy)
    cout.base.copy_(fout.base.cpu())          # identical masks on both sides
    ops.bn_bwd_reduce(dout, out, y, None, None, small[2], small[3], bs)
    sh.bn_bwd_reduce(cdout, cout, cy, None, None, csmall[2], csmall[3], cbs)
    assert rel_l2(bs, cbs) < (1e-5 if dtype == torch.float32 else 1e-3)
    assert rel_l2(fdo.base.float(), cdout.base.float()) < 1e-6
    dg, db, ds = (torch.ones(C, device=DEV) for _ in range(3))
    cdg, cdb, cds = dg.cpu().clone(), db.cpu().clone(), ds.cpu().clone()
    ops.bn_bwd_apply(dout, True, y, None, None, small[2], small[3], gamma, bs, count, add_d, dy, dg, db, ds, True)
    sh.bn_bwd_apply(cdout, True, cy, None, None, csmall[2], csmall[3], gamma.cpu(), cbs, count, cadd, cdy, cdg, cdb, cds, True)
    assert rel_l2(fdy.base.float(), cdy.base.float()) < _tol(dtype)
    assert max_rel(dg, cdg) < 1e-3 and max_rel(db, cdb) < 1e-3 and max_rel(ds, cds) < 1e-3
    # pass 1 with the max-pool backward folded in
    dpool2, _ = rand_view(N, H // 2, W // 2, C, dtype, DEV, gen=g)
    dout2, fdo2 = rand_view(N, H, W, C, dtype, DEV, ctot, c0, gen=g)
    cdout2, cdpool2 = mirror(dout2), mirror(dpool2)
    bs.zero_(); cbs.zero_()
    ops.bn_bwd_reduce(dout2, out, y, None, None, small[2], small[3], bs, dpool2)
    sh.bn_bwd_reduce(cdout2, cout, cy, None, None, csmall[2], csmall[3], cbs, cdpool2)
    assert rel_l2(bs, cbs) < (1e-5 if dtype == torch.float32 else 2e-3)
    assert rel_l2(fdo2.base.float(), cdout2.base.float()) < _tol(dtype)
    # pass 1 / 2 with the ReLU mask recomputed from y*scale+shift (no `out` tensor)
    bs.zero_(); cbs.zero_()
    ops.bn_bwd_reduce(dout, None, y, small[0], small[1], small[2], small[3], bs)
    sh.bn_bwd_reduce(cdout, None, cy, csmall[0], csmall[1], csmall[2], csmall[3], cbs)
    assert rel_l2(bs, cbs) < (1e-5 if dtype == torch.float32 else 1e-3)
    ops.bn_bwd_apply(dout, False, y, small[0], small[1], small[2], small[3], gamma, bs, count, None, dy, dg, db, None, False)
    sh.bn_bwd_apply(cdout, False, cy, csmall[0], csmall[1], csmall[2], csmall[3], gamma.cpu(), cbs, count, None, cdy, cdg, cdb, None, False)
    assert rel_l2(fdy.base.float(), cdy.base.float()) < _tol(dtype)
    assert max_rel(dg, cdg) < 1e-3 and max_rel(db, cdb) < 1e-3
    # max-pool backward (accumulate and assign) + channel sum
    dpool, _ = rand_view(N, H // 2, W // 2, C, dtype, DEV, gen=g)
    dx, fdx = rand_view(N, H, W, C, dtype, DEV, gen=g)
    cdpool, cdx = mirror(dpool), mirror(dx)
    for acc in (True, False):
        ops.maxpool2x2_bwd(out, dpool, dx, acc)
        sh.maxpool2x2_bwd(cout, cdpool, cdx, acc)
        assert rel_l2(fdx.base.float(), cdx.base.float()) < _tol(dtype)
    cs, ccs = torch.ones(C, device=DEV), torch.ones(C)
    for acc in (True, False):
        ops.channel_sum(dx, cs, acc)
        sh.channel_sum(cdx, ccs, acc)
        assert max_rel(cs, ccs) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("layout,N,H,W", [("sliced", 3, 12, 20), ("planar", 3, 12, 20), ("planar", 5, 52, 36)])
def test_ecam_forward_backward(ops, sh, dtype, layout, N, H, W):
    """sliced = channel slices of one concat buffer (register-resident kernels); planar = one dense tensor per view
    (the engine's default slot layout -> cp.async.bulk-staged kernels; 52x36 spans many 128-pixel chunks + a ragged tail)."""
    g = _gen(5)
    Cb, J, K = 32, 4, 3
    hid, hid1 = 8, 8
    CC, CT = J * Cb, (J + 1) * Cb
    if layout == "sliced":
        _, full = rand_view(N, H, W, 6 * Cb, dtype, DEV, gen=g)
        xs = [full.ch((2 + j) * Cb, Cb) for j in range(J)]
        cfull = mirror(full)
        cxs = [cfull.ch((2 + j) * Cb, Cb) for j in range(J)]
    else:
        xs = [rand_view(N, H, W, Cb, dtype, DEV, gen=g)[0] for _ in range(J)]
        cxs = [mirror(v) for v in xs]
    wts = [torch.randn(n, generator=g) * s for n, s in ((hid * CC, 0.2), (CC * hid, 0.3), (hid1 * Cb, 0.3), (Cb * hid1, 0.3), (K * CC, 0.2), (K, 0.1))]
    dw = [w.to(DEV) for w in wts]
    mk = lambda n, dt=torch.float32: (torch.zeros(n, dtype=dt, device=DEV), torch.zeros(n, dtype=dt))
    pooled, cpooled = mk(N * 2 * CT)
    argmax, cargmax = mk(N * CT, torch.int32)
    scratch = torch.zeros(N * CT, dtype=torch.int64, device=DEV)
    gates, cgates = mk(N * CT)
    hidden, chidden = mk(N * 2 * (hid + hid1))
    logits, clogits = torch.zeros(N, K, H, W, device=DEV), torch.zeros(N, K, H, W)
    ops.ecam_pool(xs, pooled, argmax, scratch)
    sh.ecam_pool(cxs, cpooled, cargmax, None)
    assert max_rel(pooled, cpooled) < (1e-5 if dtype == torch.float32 else 1e-2)
    ops.ecam_gates(N, Cb, J, hid, hid1, pooled, *dw[:4], gates, hidden)
    sh.ecam_gates(N, Cb, J, hid, hid1, pooled.cpu(), *wts[:4], cgates, chidden)
    assert max_rel(gates, cgates) < 1e-5 and max_rel(hidden, chidden) < 1e-5
    ops.ecam_final(xs, gates, dw[4], dw[5], K, logits, pooled, argmax)      # also discovers the arg-max pixels
    assert torch.equal(argmax.cpu(), cargmax)                                # both sides see the same stored values
    sh.ecam_final(cxs, gates.cpu(), wts[4], wts[5], K, clogits)
    assert rel_l2(logits, clogits) < 1e-5
    # backward
    dl = torch.randn(N, K, H, W, generator=g) * 0.01
    red, cred = mk(N * (K * CC + K), torch.float64)
    ops.ecam_bwd_reduce(xs, K, dl.to(DEV), red)
    sh.ecam_bwd_reduce(cxs, K, dl, cred)
    assert rel_l2(red, cred) < 1e-5
    dpooled, cdpooled = mk(N * 2 * CT)
    grads = [mk(w.numel()) for w in wts]
    ops.ecam_gates_bwd(N, Cb, J, hid, hid1, K, pooled, hidden, gates, red, dw[4], *dw[:4], dpooled,
                       grads[4][0], grads[5][0], grads[0][0], grads[1][0], grads[2][0], grads[3][0], False)
    sh.ecam_gates_bwd(N, Cb, J, hid, hid1, K, pooled.cpu(), hidden.cpu(), gates.cpu(), red.cpu(), wts[4], *wts[:4], cdpooled,
                      grads[4][1], grads[5][1], grads[0][1], grads[1][1], grads[2][1], grads[3][1], False)
    assert max_rel(dpooled, cdpooled) < 1e-4
    for a, b in grads:
        assert max_rel(a, b) < 1e-4
    _, dfull = rand_view(N, H, W, 6 * Cb, dtype, DEV, gen=g)
    dxs = [dfull.ch((2 + j) * Cb, Cb) for j in range(J)]
    cdfull = mirror(dfull)
    cdxs = [cdfull.ch((2 + j) * Cb, Cb) for j in range(J)]
    ops.ecam_bwd_apply(dxs, gates, dw[4], K, dl.to(DEV), dpooled, argmax)
    sh.ecam_bwd_apply(cdxs, gates.cpu(), wts[4], K, dl, dpooled.cpu(), argmax.cpu())
    assert rel_l2(dfull.base.float(), cdfull.base.float()) < _tol(dtype)


@pytest.mark.parametrize("tag", ["small", "weighted", "ragged", "allignored"])
def test_loss_kernel_vs_reference_golden(ops, golden_dir, tag):
    z = np.load(golden_dir / "loss_cases.npz")
    logits = torch.from_numpy(z[f"{tag}.logits"]).to(DEV)
    labels = torch.from_numpy(z[f"{tag}.labels"]).to(DEV)
    w = torch.from_numpy(z[f"{tag}.weights"]).to(DEV)
    N = logits.shape[0]
    loss3 = torch.zeros(3, device=DEV)
    dl = torch.zeros_like(logits)
    pred = torch.zeros(labels.shape, dtype=torch.uint8, device=DEV)
    ops.ce_dice(logits, labels, w, 3, 1.0, loss3, dl, pred, ops.ce_dice_workspace(N, DEV))
    torch.cuda.synchronize()
    assert torch.equal(pred.cpu(), torch.from_numpy(z[f"{tag}.argmax"]))           # bit-exact class map
    np.testing.assert_allclose(loss3[1].item(), z[f"{tag}.dice"], rtol=1e-5)
    if tag == "allignored":
        assert np.isnan(loss3[0].item())
        return
    np.testing.assert_allclose(loss3[0].item(), z[f"{tag}.loss"], rtol=1e-5)       # tolerance: 1e-5 rel (fp32)
    np.testing.assert_allclose(dl.cpu().numpy(), z[f"{tag}.dlogits"], rtol=1e-3, atol=1e-9)


def test_loss_kernel_full_size_vs_oracle_and_properties(ops):
    from oracle.loss_oracle import ce_dice
    g = _gen(6)
    N, H, W = 64, 224, 224                                   # BASELINE.json bs=64
    logits = (2.0 * torch.randn(N, 3, H, W, generator=g)).to(DEV)
    labels = torch.multinomial(torch.tensor([0.897, 0.024, 0.041, 0.038]), N * H * W, True, generator=g).view(N, H, W).to(DEV)
    w = torch.tensor([1.0, 1.0, 1.0], device=DEV)
    loss3, dl = torch.zeros(3, device=DEV), torch.zeros_like(logits)
    pred = torch.zeros(N, H, W, dtype=torch.uint8, device=DEV)
    ws = ops.ce_dice_workspace(N, DEV)
    ops.ce_dice(logits, labels, w, 3, 1.0, loss3, dl, pred, ws)
    assert torch.equal(pred.long(), logits.argmax(1))
    # size-independent properties: softmax gradients sum to zero over classes; forward-only call agrees; grad_scale is linear
    assert dl.sum(1).abs().max().item() < 1e-9
    loss3b = torch.zeros(3, device=DEV)
    ops.ce_dice(logits, labels, w, 3, 1.0, loss3b, None, None, ws)
    assert torch.allclose(loss3, loss3b, rtol=1e-6)
    dl2 = torch.zeros_like(dl)
    ops.ce_dice(logits, labels, w, 3, 0.5, loss3b, dl2, None, ws)
    assert torch.allclose(dl2, 0.5 * dl, rtol=1e-5, atol=1e-12)
    # oracle on a 4-sample slice (per-sample dice terms make the loss separable only for the CE part: use N=4 run)
    sub = slice(0, 4)
    ops.ce_dice(logits[sub].contiguous(), labels[sub].contiguous(), w, 3, 1.0, loss3b, dl2[sub], None, ops.ce_dice_workspace(4, DEV))
    r = ce_dice(logits[sub].cpu().numpy(), labels[sub].cpu().numpy(), [1.0, 1.0, 1.0], 3)
    np.testing.assert_allclose(loss3b[0].item(), r["loss"], rtol=1e-5)
    np.testing.assert_allclose(dl2[sub].cpu().numpy(), r["dlogits"], rtol=2e-3, atol=1e-10)


@pytest.mark.parametrize("N,H,W", [(3, 64, 64), (5, 224, 224), (64, 224, 224), (2, 36, 100), (150, 64, 64), (70, 224, 224)])
def test_loss_resident_single_pass_matches_two_pass_and_oracle(ops, N, H, W):
    """The resident single-pass CE+Dice kernel (logits kept on the SMs across the grid barrier) against the two-pass kernels on the same
    workspace (alternating calls exercise the self-cleaning flip protocol) and, on a 2-sample slice, the numpy oracle.  Shapes: whole and
    ragged 2048-pixel chunks, more chunks than CTAs, fewer chunks than CTAs, and (70 x 224 x 224) beyond the on-chip capacity (falls back)."""
    from oracle.loss_oracle import ce_dice
    g = _gen(60 + N)
    logits = (2.0 * torch.randn(N, 3, H, W, generator=g)).to(DEV)
    labels = torch.randint(0, 4, (N, H, W), generator=g).to(DEV)
    w = torch.tensor([0.5, 2.0, 1.25], device=DEV)
    ws = ops.ce_dice_workspace(N, DEV)
    outs = []
    try:
        for variant in (0, 1, 0, 0, 1):
            ops.set_option("loss_variant", variant)
            loss3, dl = torch.zeros(3, device=DEV), torch.full_like(logits, 7.0)
            pred = torch.full((N, H, W), 9, dtype=torch.uint8, device=DEV)
            ops.ce_dice(logits, labels, w, 3, 1.0, loss3, dl, pred, ws, dice_weight=0.7)
            torch.cuda.synchronize()
            outs.append((loss3.clone(), dl, pred))
    finally:
        ops.set_option("loss_variant", 0)
    l0, d0, p0 = outs[0]
    assert torch.equal(p0.long(), logits.argmax(1))
    for l, d, pr in outs[1:]:
        assert torch.equal(pr, p0)
        assert torch.allclose(l, l0, rtol=5e-6), (l, l0)
        assert torch.allclose(d, d0, rtol=1e-4, atol=1e-10)      # |d| ~ 1 / (N H W) = 1e-4 .. 1e-6; cancellation in w_y (p_y - 1)
    assert torch.equal(outs[2][1], d0) and torch.equal(outs[3][1], d0)          # same kernel, same workspace state -> bit-identical
    sub = slice(0, 2)
    loss3b, dl2 = torch.zeros(3, device=DEV), torch.zeros_like(logits[sub])
    ops.ce_dice(logits[sub].contiguous(), labels[sub].contiguous(), w, 3, 1.0, loss3b, dl2, None, ops.ce_dice_workspace(2, DEV))
    r = ce_dice(logits[sub].cpu().numpy(), labels[sub].cpu().numpy(), [0.5, 2.0, 1.25], 3)
    np.testing.assert_allclose(loss3b[0].item(), r["loss"], rtol=1e-5)
    np.testing.assert_allclose(dl2.cpu().numpy(), r["dlogits"], rtol=2e-3, atol=1e-10)


def test_adam(ops, sh):
    g = _gen(7)
    n = 10007
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g)
    m, v = torch.zeros(n), torch.zeros(n)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    dp, dg, dm, dv = (t.clone().to(DEV) for t in (p, gr, m, v))
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    for it in range(3):
        ref.grad = gr.clone() * (it + 1)
        opt.step()
        ops.adam_step(dp, dg * (it + 1), dm, dv, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, step)
    assert int(step.item()) == 3
    assert max_rel(dp, ref.data) < 1e-6


def test_adamw(ops, sh):
    """Fused AdamW (change_detection_trainer.py:55-60: betas + decoupled weight decay) vs torch.optim.AdamW and the shadow."""
    g = _gen(8)
    n = 10007
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.8, 0.95), weight_decay=0.05)
    dp, dg = p.clone().to(DEV), gr.clone().to(DEV)
    dm, dv = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    sp, sm, sv, sstep = p.clone(), torch.zeros(n), torch.zeros(n), torch.zeros(1, dtype=torch.int32)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    for it in range(4):
        ref.grad = gr.clone() * (it + 1)
        opt.step()
        ops.adamw_step(dp, dg * (it + 1), dm, dv, 3e-3, 0.8, 0.95, 1e-8, 0.05, 1.0, step)
        sh.adamw_step(sp, gr * (it + 1), sm, sv, 3e-3, 0.8, 0.95, 1e-8, 0.05, 1.0, sstep)
    assert int(step.item()) == 4
    assert max_rel(dp, ref.data) < 1e-6 and max_rel(sp, ref.data) < 1e-6


def test_confusion_update(ops, sh):
    """Bit-exact integer work: the fused confusion-matrix kernel against torch.bincount, incl. accumulation over two calls."""
    g = _gen(9)
    n = 3 * 224 * 224 + 5
    pred = torch.randint(0, 3, (n,), generator=g, dtype=torch.uint8)
    lab = torch.randint(0, 4, (n,), generator=g, dtype=torch.int64)
    mat, cmat = torch.zeros(4, 4, dtype=torch.int64, device=DEV), torch.zeros(4, 4, dtype=torch.int64)
    for _ in range(2):
        ops.confusion_update(pred.to(DEV), lab.to(DEV), 4, 3, mat)
        sh.confusion_update(pred, lab, 4, 3, cmat)
    assert torch.equal(mat.cpu(), cmat) and int(cmat[3].sum()) == 0
    from kurosiwo_b200.utilities import ConfusionMetrics
    m = ConfusionMetrics(3, 3, DEV)
    m.update(pred.to(DEV), lab.to(DEV))
    m2 = ConfusionMetrics(3, 3, "cpu")
    m2.update(pred, lab)
    for a, b in zip(m.compute(), m2.compute()):
        assert torch.allclose(a.cpu(), b)


@pytest.mark.parametrize("weights", [[1.0, 1.0, 1.0], [0.3715753140309927, 14.009780283125977, 8.20405370357821]])
def test_cross_entropy_mode_matches_torch(ops, weights):
    """dice_weight = 0: the fused loss kernel is exactly nn.CrossEntropyLoss(weight, ignore_index=3) (the reference's default
    criterion, utilities/utilities.py:308-321): value 1e-5, gradient 1e-5 rel-L2, argmax bit-exact."""
    import torch.nn.functional as F
    from kurosiwo_b200.bce_and_dice import FusedCrossEntropyLoss
    g = _gen(11)
    N, H, W = 3, 28, 36
    z = (3.0 * torch.randn(N, 3, H, W, generator=g)).requires_grad_(True)
    y = torch.randint(0, 4, (N, H, W), generator=g)
    w = torch.tensor(weights)
    ref = F.cross_entropy(z, y, weight=w, ignore_index=3)
    ref.backward()
    zd = z.detach().to(DEV).requires_grad_(True)
    crit = FusedCrossEntropyLoss(weight=w, ignore_index=3).to(DEV)
    loss = crit(zd, y.to(DEV))
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-5)
    assert rel_l2(zd.grad, z.grad) < 1e-5
    assert torch.equal(crit.last_pred.cpu().long(), z.detach().argmax(1))


def test_confusion_update_grouped(ops, sh):
    """Bit-exact: ONE launch fills the global, the per-activation and the per-climate-zone matrices (reference
    change_detection_trainer.py:184-199, :445-480) - against the shadow (torch.bincount per sample and group)."""
    g = _gen(10)
    B, H, W = 7, 96, 112
    pred = torch.randint(0, 3, (B, H, W), generator=g, dtype=torch.uint8)
    lab = torch.randint(0, 4, (B, H, W), generator=g, dtype=torch.int64)
    ka = torch.tensor([0, 2, 2, -1, 4, 0, 9], dtype=torch.int32)     # -1 and 9 (>= n_a) are skipped
    kb = torch.tensor([0, 1, 2, 2, 1, 0, 0], dtype=torch.int32)
    mats = [torch.zeros(4, 4, dtype=torch.int64), torch.zeros(5, 4, 4, dtype=torch.int64), torch.zeros(3, 4, 4, dtype=torch.int64)]
    dmats = [m.clone().to(DEV) for m in mats]
    for _ in range(2):
        ops.confusion_update_grouped(pred.to(DEV), lab.to(DEV), 4, 3, dmats[0], ka.to(DEV), dmats[1], kb.to(DEV), dmats[2])
        sh.confusion_update_grouped(pred, lab, 4, 3, mats[0], ka, mats[1], kb, mats[2])
    for d, m in zip(dmats, mats):
        assert torch.equal(d.cpu(), m)
    assert int(mats[1].sum()) == 2 * int((lab[[0, 1, 2, 4, 5]] != 3).sum()) and int(mats[2].sum()) == int(mats[0].sum())
    from kurosiwo_b200.utilities import GroupedConfusionMetrics
    acts = [130, 470, 555]
    m = GroupedConfusionMetrics(3, 3, DEV, activations=acts, zones=True)
    activ = torch.tensor([470, 130, 470, 555, 130, 7, 555])
    clz = torch.tensor([1, 2, 2, 3, 1, 2, 3])
    m.update(pred.to(DEV), lab.to(DEV), activ=activ, clz=clz)
    c = GroupedConfusionMetrics(3, 3, "cpu", activations=acts, zones=True)
    c.update(pred, lab, activ=activ, clz=clz)
    assert torch.equal(m.mat.cpu(), c.mat) and torch.equal(m.mat_aoi.cpu(), c.mat_aoi) and torch.equal(m.mat_zone.cpu(), c.mat_zone)


def test_sar_preprocess_matches_reference_transform(ops, sh):
    """Input pipeline on the GPU (reference dataset/Dataset.py:162-168 clamp + nan_to_num, :192-198 Normalize) - BIT-exact against
    the reference's own torch ops on seeded raw tiles with NaN / +-inf / negative / over-range values, both clamp modes, in place."""
    g = _gen(12)
    mean, std = torch.tensor([0.0953, 0.0264]), torch.tensor([0.0427, 0.0215])
    for (B, H, W) in ((3, 224, 224), (2, 7, 9)):
        raw = torch.empty(B, 2, H, W).exponential_(1.0, generator=g) * mean.view(1, 2, 1, 1) * 2 - 0.01
        flat = raw.view(-1)
        idx = torch.randperm(flat.numel(), generator=g)[:64]
        flat[idx[:24]] = float("nan"); flat[idx[24:40]] = float("inf"); flat[idx[40:56]] = float("-inf"); flat[idx[56:]] = 1e30
        for clamp in (0.15, 0.0):
            want = torch.empty_like(raw)
            sh.sar_preprocess(raw, want, mean, std, clamp)
            x = raw.clone().to(DEV)
            ops.sar_preprocess(x, x, mean.to(DEV), std.to(DEV), clamp)
            assert torch.equal(x.cpu(), want), (B, H, W, clamp, float((x.cpu() - want).abs().max()))
    # through the trainer surface: raw tiles in, the same tensors the normalised loader would have produced
    from kurosiwo_b200.change_detection_trainer import preprocess_raw
    cfg = {"raw_input": True, "data_mean": [0.0953, 0.0264], "data_std": [0.0427, 0.0215], "clamp_input": 0.15}
    raw = torch.empty(2, 2, 32, 32).exponential_(1.0, generator=g) * mean.view(1, 2, 1, 1)
    want = (raw.clamp(0, 0.15) - mean.view(1, 2, 1, 1)) / std.view(1, 2, 1, 1)
    assert torch.equal(preprocess_raw(raw.to(DEV), cfg).cpu(), want)
