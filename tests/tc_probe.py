"""Stand-alone probe of the tcgen05 conv / wgrad kernels against the CUDA-core kernels (same inputs).

Run on a B200:  python tests/tc_probe.py [out.json]
Each case is tried for every (bo_mode, MT) so that a descriptor-convention problem shows up as a
pattern instead of a single failure.  Used by tests/test_gpu_tc.py through a subprocess + timeout so a
hung kernel cannot hang the test session.
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from kurosiwo_b200.lib import IMPL_SIMT, IMPL_TC, CudaOps, KsError, View  # noqa: E402
from gpu_util import mirror, rand_view, rel_l2  # noqa: E402
from shadow_ops import ShadowOps  # noqa: E402

dev = "cuda:0"
bf = torch.bfloat16


def conv_case(ops, name, N, H, W, ks, src_specs, dst_specs, bias, gen):
    """src_specs/dst_specs: list of (C, ctot, c0, kind) kind in {'plain','phases'}; returns dict of errors."""
    srcs, dsts, dst_fulls = [], [], []
    for (C, ctot, c0, kind) in src_specs:
        if kind == "phases":
            _, full = rand_view(N, 2 * H, 2 * W, C, bf, dev, gen=gen)
            srcs += [full.phase(k // 2, k % 2) for k in range(4)]
        else:
            v, _ = rand_view(N, H, W, C, bf, dev, ctot, c0, gen=gen)
            srcs.append(v)
    accs = []
    for (C, ctot, c0, kind, acc) in dst_specs:
        if kind == "phases":
            _, full = rand_view(N, 2 * H, 2 * W, C, bf, dev, gen=gen)
            dsts += [full.phase(k // 2, k % 2) for k in range(4)]
            accs += [acc] * 4
            dst_fulls.append(full)
        else:
            v, full = rand_view(N, H, W, C, bf, dev, ctot, c0, gen=gen)
            dsts.append(v)
            accs.append(acc)
            dst_fulls.append(full)
    cin, cout = sum(v.C for v in srcs), sum(v.C for v in dsts)
    w = (torch.randn(ks * ks * cout * cin, generator=gen) * (2.0 / (cin * ks * ks)) ** 0.5).to(bf).to(dev)
    b = (torch.randn(cout, generator=gen) * 0.1).to(dev) if bias else None
    init = [f.base.clone() for f in dst_fulls]
    # reference 1: the torch-CPU shadow of the op (fp64 accumulate over the same bf16 inputs), independent of every CUDA kernel
    sh = ShadowOps()
    m_srcs = [mirror(v) for v in srcs]
    m_fulls = [mirror(f) for f in dst_fulls]
    m_dsts = []
    for (C, ctot, c0, kind, acc), mf in zip(dst_specs, m_fulls):
        m_dsts += [mf.phase(k // 2, k % 2) for k in range(4)] if kind == "phases" else [mf.ch(c0, C)]
    sh.conv2d(N, H, W, ks, m_srcs, w.cpu(), b.cpu() if b is not None else None, m_dsts, accs, None)
    shadow = [mf.base.clone() for mf in m_fulls]
    # reference 2: CUDA-core kernel
    ops.conv2d(N, H, W, ks, srcs, w, b, dsts, accs, None, IMPL_SIMT)
    torch.cuda.synchronize()
    ref = [f.base.clone() for f in dst_fulls]
    want_stats = len(dsts) == 1 and not accs[0]
    ref_stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    if want_stats:
        ops.bn_stats(dsts[0], ref_stats)
    res = {}
    variants = {
        "v1_mt1": {"tc_v1": 1, "tc_mt": 1},
        "v2_auto": {},
        "v2_res_mt2": {"tc_mt": 2},
        "v2_nores_mt1": {"tc_no_resident": 1, "tc_mt": 1},
        "v2_nores_mt2": {"tc_no_resident": 1, "tc_mt": 2},
        "v2_nores_mt4": {"tc_no_resident": 1, "tc_mt": 4},
        "v2_no_ns3": {"tc_no_ns3": 1},
        "v2_no_ns3_nores": {"tc_no_ns3": 1, "tc_no_resident": 1},
        # column taps stacked along N for every 3x3 case with N tile <= 80 (the library default applies it from Cin >= 96)
        "v2_ns3": {"tc_ns3_min_cin": 1},
        "v2_ns3_nores_one_cta": {"tc_ns3_min_cin": 1, "tc_no_resident": 1, "tc_ns3_mode": 2},
        "v2_ns3_nores_two_cta": {"tc_ns3_min_cin": 1, "tc_no_resident": 1, "tc_ns3_mode": 1},
        "v2_ns3_nores_mt2": {"tc_ns3_min_cin": 1, "tc_no_resident": 1, "tc_ns3_mode": 2, "tc_mt": 2},
        # 16 epilogue warps (576 threads, one CTA per SM)
        "v2_ew16": {"tc_ew": 16},
        "v2_ew16_ns3": {"tc_ew": 16, "tc_ns3_min_cin": 1},
        "v2_ew16_nores_mt4": {"tc_ew": 16, "tc_no_resident": 1, "tc_mt": 4},
        # BatchNorm statistics: the shuffle butterfly everywhere (the default keeps them in registers for N tiles <= 32), also stacked
        "v2_stat_butterfly": {"tc_stat_mode": 1},
        "v2_stat_butterfly_ns3": {"tc_stat_mode": 1, "tc_ns3_min_cin": 1},
    }
    OPTS = ("tc_v1", "tc_mt", "tc_no_resident", "tc_no_ns3", "tc_ns3_min_cin", "tc_ns3_mode", "tc_ew", "tc_stat_mode")
    for key, opts in variants.items():
        for f, i0 in zip(dst_fulls, init):
            f.base.copy_(i0)
        for o in OPTS:
            ops.set_option(o, opts.get(o, 0))
        try:
            st = torch.zeros(2 * cout, dtype=torch.float64, device=dev) if want_stats else None
            ops.conv2d(N, H, W, ks, srcs, w, b, dsts, accs, st, IMPL_TC)
            torch.cuda.synchronize()
            err = max(rel_l2(f.base.float(), r.float()) for f, r in zip(dst_fulls, ref))
            err = max(err, max(rel_l2(f.base.float(), r.float()) for f, r in zip(dst_fulls, shadow)))
            if want_stats:
                err = max(err, rel_l2(st, ref_stats))
            res[key] = err
        except KsError as e:
            res[key] = f"error: {e}"
    for o in OPTS:
        ops.set_option(o, 0)
    return res


def wgrad_case(ops, name, N, H, W, ks, x_specs, dy_specs, gen):
    xs, dys = [], []
    for specs, out in ((x_specs, xs), (dy_specs, dys)):
        for (C, ctot, c0, kind) in specs:
            if kind == "phases":
                _, full = rand_view(N, 2 * H, 2 * W, C, bf, dev, gen=gen)
                out += [full.phase(k // 2, k % 2) for k in range(4)]
            else:
                v, _ = rand_view(N, H, W, C, bf, dev, ctot, c0, gen=gen)
                out.append(v)
    cin, cout = sum(v.C for v in xs), sum(v.C for v in dys)
    ref = torch.zeros(ks * ks * cout * cin, device=dev)
    got = torch.zeros_like(ref)
    ops.conv2d_wgrad(N, H, W, ks, xs, dys, ref, False, IMPL_SIMT)
    torch.cuda.synchronize()
    shadow = torch.zeros(ks * ks * cout * cin)
    ShadowOps().conv2d_wgrad(N, H, W, ks, [mirror(v) for v in xs], [mirror(v) for v in dys], shadow, False)
    res = {"simt_vs_shadow": rel_l2(ref, shadow)}
    for mode, tag in ((0, ""), (2, "_halo")):
        ops.set_option("wgrad_mode", mode)
        try:
            ops.conv2d_wgrad(N, H, W, ks, xs, dys, got, False, IMPL_TC)
            torch.cuda.synchronize()
            res["assign" + tag] = max(rel_l2(got, ref), rel_l2(got, shadow))
            ops.conv2d_wgrad(N, H, W, ks, xs, dys, got, True, IMPL_TC)  # accumulate: 2x
            torch.cuda.synchronize()
            res["accumulate" + tag] = rel_l2(got, 2 * ref)
        except KsError as e:
            res["error" + tag] = str(e)
    ops.set_option("wgrad_mode", 0)
    return res


def main():
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else None
    ops = CudaOps()
    gen = torch.Generator().manual_seed(1234)
    report = {"conv": {}, "wgrad": {}}
    P = "plain"
    conv_cases = {
        "A_k3_c64_n32": (2, 16, 28, 3, [(64, 64, 0, P)], [(32, 32, 0, P, False)], True),
        "B_k3_prefix96+64_n32_bk32": (2, 16, 16, 3, [(96, 192, 0, P), (64, 64, 0, P)], [(32, 192, 64, P, False)], True),
        "C_k3_c128_n256": (3, 8, 8, 3, [(128, 128, 0, P)], [(256, 256, 0, P, False)], True),
        "D_k3_h14_c256_n512": (2, 14, 14, 3, [(256, 256, 0, P)], [(512, 512, 0, P, False)], True),
        "E_k1_convT_fwd": (2, 8, 8, 1, [(64, 64, 0, P)], [(64, 64, 0, "phases", False)], True),
        "F_k1_convT_dgrad_acc": (2, 8, 8, 1, [(64, 64, 0, "phases")], [(64, 128, 64, P, True)], False),
        "G_k3_dgrad_multidst": (2, 16, 28, 3, [(32, 32, 0, P)], [(64, 192, 0, P, True), (32, 192, 64, P, False), (64, 64, 0, P, False)], False),
        "H_k3_odd_hw": (1, 20, 20, 3, [(64, 64, 0, P)], [(64, 64, 0, P, False)], True),
        "I_k3_big": (4, 56, 56, 3, [(128, 256, 0, P), (128, 128, 0, P)], [(128, 128, 0, P, False)], True),
        # column-tap-stacked N (narrow N tiles): resident (N = 32 / 16) and streamed (N = 64) weights, accumulate, ragged edges
        "J_k3_c64_n64": (2, 28, 28, 3, [(64, 64, 0, P)], [(64, 64, 0, P, False)], True),
        "K_k3_c320_n64": (2, 28, 28, 3, [(256, 256, 0, P), (64, 128, 64, P)], [(64, 128, 0, P, False)], True),
        "L_k3_c32_n32_acc": (3, 16, 28, 3, [(32, 32, 0, P)], [(32, 96, 64, P, True)], False),
        "M_k3_c224_n32_bk32": (2, 24, 42, 3, [(160, 160, 0, P), (64, 64, 0, P)], [(32, 32, 0, P, False)], True),
        "N_k3_c64_n16": (2, 16, 14, 3, [(64, 64, 0, P)], [(16, 16, 0, P, False)], True),
        "O_k3_c64_n64_ragged": (5, 21, 30, 3, [(64, 64, 0, P)], [(64, 64, 0, P, False)], True),
        "P_k3_c128_n80": (2, 14, 14, 3, [(128, 128, 0, P)], [(80, 80, 0, P, False)], False),
    }
    for name, (N, H, W, ks, ss, ds, bias) in conv_cases.items():
        report["conv"][name] = conv_case(ops, name, N, H, W, ks, ss, ds, bias, gen)
        print(name, report["conv"][name], flush=True)
    wgrad_cases = {
        "H_k3_x96+64_dy32": (2, 16, 32, 3, [(96, 192, 0, P), (64, 64, 0, P)], [(32, 32, 0, P)]),
        "I_k3_x128_dy128": (2, 16, 16, 3, [(128, 128, 0, P)], [(128, 128, 0, P)]),
        "J_k3_x256_dy256": (2, 8, 8, 3, [(256, 256, 0, P)], [(256, 256, 0, P)]),
        "K_k1_convT": (2, 8, 8, 1, [(64, 64, 0, P)], [(64, 64, 0, "phases")]),
        "L_k3_x64_dy64_odd": (3, 20, 12, 3, [(64, 64, 0, P)], [(64, 64, 0, P)]),
        "M_k3_big": (4, 56, 56, 3, [(192, 256, 0, P), (128, 128, 0, P)], [(64, 64, 0, P)]),
        "N_k3_x32_dy32_w28": (3, 16, 28, 3, [(32, 32, 0, P)], [(32, 32, 0, P)]),
        "O_k3_x160_dy32_w28": (2, 24, 28, 3, [(96, 96, 0, P), (64, 64, 0, P)], [(32, 96, 32, P)]),
        "P_k3_x128_dy128_w14": (4, 14, 14, 3, [(128, 128, 0, P)], [(128, 128, 0, P)]),
        "Q_k3_x64_dy64_w56": (2, 56, 56, 3, [(64, 192, 64, P)], [(64, 64, 0, P)]),
    }
    for name, (N, H, W, ks, xs, ys) in wgrad_cases.items():
        report["wgrad"][name] = wgrad_case(ops, name, N, H, W, ks, xs, ys, gen)
        print(name, report["wgrad"][name], flush=True)
    if out:
        out.parent.mkdir(parents=True, exist_ok=True)
        out.write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
