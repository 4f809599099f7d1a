"""FloodViT + UPerNet head (BASELINE.json configs[3]: "MAE-ViT-B encoder + UPerNet head") on the C-ABI ops.

Head semantics: HF transformers `UperNetHead` (modeling_upernet.py; the third-party code the reference reaches through
models/upernet.py:80, transformers==4.31.0 at requirements.txt:27): every ConvModule = conv(bias=False) -> BatchNorm2d(train) -> ReLU;
lateral 1x1 modules on the first three features, pyramid pooling (AdaptiveAvgPool2d 1/2/3/6 -> 1x1 module -> bilinear) + 3x3
bottleneck on the last, top-down adds, 3x3 FPN modules, concat, 3x3 fpn_bottleneck, 1x1 classifier, logits resized to the input.
The reference has no ViT + UPerNet composition; defined here as: token maps (cls dropped) after the ViT
blocks `out_indices` (default depth/4, depth/2, 3depth/4, depth - the last one after the final LayerNorm), all G x G, so the head's
resizes between levels are identities.  The feature maps are strided VIEWS of the token buffers (no copies); their gradients enter the
encoder's residual-stream gradient at the matching depth through the lateral convs' data-gradient epilogue (+=).
"""
from __future__ import annotations

from typing import List

import torch

from .lib import View
from .vit_engine import ViTSegEngine, tok_view

BN_EPS, BN_MOMENTUM = 1e-5, 0.1
SCALES = (1, 2, 3, 6)
CLS_PAD = 32


class _CBR:
    """One ConvModule execution: pre-BN conv output y, ReLU output out, their gradients, BN scratch."""

    def __init__(self, eng, name, cin, cout, k, n, h, w):
        z = lambda c: torch.zeros(n * h * w, c, dtype=eng.dtype, device=eng.device)
        self.name, self.cin, self.cout, self.k, self.n, self.h, self.w = name, cin, cout, k, n, h, w
        self.y, self.out, self.dout, self.dy = z(cout), z(cout), z(cout), z(cout)
        self.bn = torch.zeros(4 * cout, dtype=torch.float32, device=eng.device)
        self.stats = torch.zeros(2 * cout, dtype=torch.float64, device=eng.device)
        self.bstats = torch.zeros(2 * cout, dtype=torch.float64, device=eng.device)
        kk = k * k
        eng.wp[f"{name}.fwd"] = torch.zeros(kk * cout * cin, dtype=eng.dtype, device=eng.device)
        eng.wp[f"{name}.dgrad"] = torch.zeros(kk * cout * cin, dtype=eng.dtype, device=eng.device)
        eng.gp[name] = torch.zeros(kk * cout * cin, dtype=torch.float32, device=eng.device)

    def v(self, t) -> View:
        Cn = t.shape[1]
        return View(t.view(-1), 0, self.n, self.h, self.w, Cn, self.h * self.w * Cn, self.w * Cn, Cn)


class ViTUperNetEngine(ViTSegEngine):
    def __init__(self, ops, module, enc_prefix, cfg, num_classes, B, H, W, dtype, device, out_indices=None, hidden=512, head_prefix="decode_head."):
        self.hp_, self.Eh = head_prefix, hidden
        depth = cfg["depth"]
        self.out_indices = list(out_indices) if out_indices else [max(1, depth // 4), max(1, depth // 2), max(1, 3 * depth // 4), depth]
        assert len(self.out_indices) == 4 and self.out_indices[-1] == depth and sorted(set(self.out_indices)) == self.out_indices
        self.gp = {}
        super().__init__(ops, module, enc_prefix, cfg, "upernet", num_classes, B, H, W, dtype, device)

    # ------------------------------------------------------------------------------------------
    def _alloc(self):
        super()._alloc()
        B, G, D, E, dev, hp = self.B, self.G, self.D, self.Eh, self.device, self.hp_
        self.lat = [_CBR(self, f"{hp}lateral_convs.{i}", D, E, 1, B, G, G) for i in range(3)]
        self.psp = [_CBR(self, f"{hp}psp_modules.{i}.1", D, E, 1, B, s, s) for i, s in enumerate(SCALES)]
        self.pooled = [torch.zeros(B * s * s, D, dtype=self.dtype, device=dev) for s in SCALES]
        self.dpooled = [torch.zeros(B * s * s, D, dtype=self.dtype, device=dev) for s in SCALES]
        self.up = [torch.zeros(B * G * G, E, dtype=self.dtype, device=dev) for _ in SCALES]
        self.dup = [torch.zeros(B * G * G, E, dtype=self.dtype, device=dev) for _ in SCALES]
        self.bott = _CBR(self, f"{hp}bottleneck", D + len(SCALES) * E, E, 3, B, G, G)
        self.td = [torch.zeros(B * G * G, E, dtype=self.dtype, device=dev) for _ in range(3)]     # top-down sums of levels 0..2
        self.dtd = [torch.zeros(B * G * G, E, dtype=self.dtype, device=dev) for _ in range(3)]
        self.fpn = [_CBR(self, f"{hp}fpn_convs.{i}", E, E, 3, B, G, G) for i in range(3)]
        self.fb = _CBR(self, f"{hp}fpn_bottleneck", 4 * E, E, 3, B, G, G)
        self.zc = torch.zeros(B * G * G, CLS_PAD, dtype=self.dtype, device=dev)
        self.dzc = torch.zeros(B * G * G, CLS_PAD, dtype=self.dtype, device=dev)
        self.wp["cls.fwd"] = torch.zeros(CLS_PAD * E, dtype=self.dtype, device=dev)
        self.wp["cls.dgrad"] = torch.zeros(E * CLS_PAD, dtype=self.dtype, device=dev)
        self.wp["cls.bias"] = torch.zeros(CLS_PAD, dtype=torch.float32, device=dev)
        self.gp["cls"] = torch.zeros(CLS_PAD * E, dtype=torch.float32, device=dev)
        self.gp["cls.bias"] = torch.zeros(CLS_PAD, dtype=torch.float32, device=dev)
        self.cbrs: List[_CBR] = self.lat + self.psp + [self.bott] + self.fpn + [self.fb]

    def _pack_jobs(self):
        P, jobs = self.params, []
        for name, (o, i) in self.linears.items():
            w = P.p(f"{name}.weight")
            jobs.append((w, self.wp[f"{name}.fwd"], (o * i,), (1,), 0))
            jobs.append((w, self.wp[f"{name}.dgrad"], (i, o), (1, i), 0))
        for L in self.cbrs:
            w, kk, co, ci = P.p(f"{L.name}.conv.weight"), L.k * L.k, L.cout, L.cin
            jobs.append((w, self.wp[f"{L.name}.fwd"], (kk, co, ci), (1, ci * kk, kk), 0))                    # [t][o][i] = w[o][i][t]
            jobs.append((w, self.wp[f"{L.name}.dgrad"], (kk, ci, co), (-1, kk, ci * kk), kk - 1))            # [t][i][o] = w[o][i][kk-1-t]
        E, K, hp = self.Eh, self.K, self.hp_
        w = P.p(f"{hp}classifier.weight")
        jobs.append((w, self.wp["cls.fwd"], (K * E,), (1,), 0))
        jobs.append((w, self.wp["cls.dgrad"], (E, K), (1, E), 0, (CLS_PAD, 1), 0))
        jobs.append((P.p(f"{hp}classifier.bias"), self.wp["cls.bias"], (K,), (1,), 0))
        return jobs

    def _unpack_jobs(self):
        P, jobs = self.params, []
        for L in self.cbrs:
            kk, co, ci = L.k * L.k, L.cout, L.cin
            jobs.append((self.gp[L.name], P.g(f"{L.name}.conv.weight"), (co, ci, kk), (ci, 1, co * ci), 0))
        jobs.append((self.gp["cls"], P.g(f"{self.hp_}classifier.weight"), (self.K * self.Eh,), (1,), 0))
        jobs.append((self.gp["cls.bias"], P.g(f"{self.hp_}classifier.bias"), (self.K,), (1,), 0))
        return jobs

    def _ensure_nbt(self):
        mods = [self.module.get_submodule(f"{L.name}.batch_norm") for L in self.cbrs]
        flat = getattr(self, "nbt_all", None)
        if flat is not None and flat.device == self.device and all(m.num_batches_tracked.data_ptr() == flat.data_ptr() + 8 * i for i, m in enumerate(mods)):
            return
        flat = torch.zeros(len(mods), dtype=torch.int64, device=self.device)
        for i, m in enumerate(mods):
            flat[i] = m.num_batches_tracked.to(self.device)
            m._buffers["num_batches_tracked"] = flat[i]
        self.nbt_all = flat

    # ------------------------------------------------------------------------------------------
    def _feat(self, k: int) -> View:
        c = self.out_indices[k]
        t = self.tok if c == self.depth else self.blocks[c].xa        # residual stream after block c = input buffer of block c+1
        return self.grid_view(t)

    def _cbr_fwd(self, L: _CBR, srcs: List[View], training: bool):
        ops, P = self.ops, self.params
        c = L.cout
        sc, sh, mu, rs = [L.bn[i * c:(i + 1) * c] for i in range(4)]
        if training:
            ops.zero_(L.stats)
        ops.conv2d(L.n, L.h, L.w, L.k, srcs, self.wp[f"{L.name}.fwd"], None, [L.v(L.y)], None, L.stats if training else None, self.conv_impl)
        m = self.module.get_submodule(f"{L.name}.batch_norm")
        if training:
            ops.bn_finalize(c, float(L.n * L.h * L.w), L.stats, P.p(f"{L.name}.batch_norm.weight"), P.p(f"{L.name}.batch_norm.bias"), BN_EPS,
                            BN_MOMENTUM, m.running_mean, m.running_var, sc, sh, mu, rs)
        else:
            torch.mul(P.p(f"{L.name}.batch_norm.weight"), torch.rsqrt(m.running_var + BN_EPS), out=sc)
            torch.sub(P.p(f"{L.name}.batch_norm.bias"), m.running_mean * sc, out=sh)
        ops.bn_act(L.v(L.y), sc, sh, None, True, L.v(L.out), None)

    def _cbr_bwd(self, L: _CBR, srcs: List[View], dout: torch.Tensor, gdsts=None, gacc=None):
        """dout: gradient of the module's ReLU output (masked in place).  Weight gradient always; data gradient into gdsts if given."""
        ops, P = self.ops, self.params
        c = L.cout
        mu, rs = L.bn[2 * c:3 * c], L.bn[3 * c:4 * c]
        ops.zero_(L.bstats)
        ops.bn_bwd_reduce(L.v(dout), L.v(L.out), L.v(L.y), None, None, mu, rs, L.bstats)
        ops.bn_bwd_apply(L.v(dout), True, L.v(L.y), None, None, mu, rs, P.p(f"{L.name}.batch_norm.weight"), L.bstats, float(L.n * L.h * L.w), None,
                         L.v(L.dy), P.g(f"{L.name}.batch_norm.weight"), P.g(f"{L.name}.batch_norm.bias"), None, False)
        ops.conv2d_wgrad(L.n, L.h, L.w, L.k, srcs, [L.v(L.dy)], self.gp[L.name], False, self.conv_impl)
        if gdsts is not None:
            ops.conv2d(L.n, L.h, L.w, L.k, [L.v(L.dy)], self.wp[f"{L.name}.dgrad"], None, gdsts, gacc, None, self.conv_impl)

    def _head_forward(self, training: bool) -> torch.Tensor:
        ops, B, G = self.ops, self.B, self.G
        if training:
            self._ensure_nbt()
            self.nbt_all.add_(1)
        feats = [self._feat(k) for k in range(4)]
        for k in range(3):
            self._cbr_fwd(self.lat[k], [feats[k]], training)
        for i, s in enumerate(SCALES):
            ops.adaptive_avgpool_fwd(feats[3], s, self.pooled[i])
            L = self.psp[i]
            self._cbr_fwd(L, [L.v(self.pooled[i])], training)
            ops.bilinear_nhwc_fwd(B, s, s, G, G, L.out, self.up[i], False)
        bv = self.bott.v
        self._cbr_fwd(self.bott, [feats[3]] + [bv(u) for u in self.up], training)
        prev = self.bott.out                                    # top-down: laterals[i-1] += resize(laterals[i]) (identity resize)
        for k in (2, 1, 0):
            self.td[k].copy_(self.lat[k].out)
            ops.bilinear_nhwc_fwd(B, G, G, G, G, prev, self.td[k], True)
            prev = self.td[k]
        for k in range(3):
            self._cbr_fwd(self.fpn[k], [bv(self.td[k])], training)
        self._cbr_fwd(self.fb, [bv(self.fpn[0].out), bv(self.fpn[1].out), bv(self.fpn[2].out), bv(self.bott.out)], training)
        ops.conv2d(B, G, G, 1, [bv(self.fb.out)], self.wp["cls.fwd"], self.wp["cls.bias"], [bv(self.zc)], None, None, self.conv_impl)
        ops.bilinear_up_fwd(B, G, G * G, 0, self.K, self.H, self.W, self.zc, self.logits)
        return self.logits

    def _head_backward(self, dlogits: torch.Tensor):
        ops, P, B, G, E = self.ops, self.params, self.B, self.G, self.Eh
        bv = self.bott.v
        feats = [self._feat(k) for k in range(4)]
        ops.bilinear_up_bwd(B, G, G * G, 0, self.K, self.H, self.W, dlogits, self.dzc)
        ops.conv2d_wgrad(B, G, G, 1, [bv(self.fb.out)], [bv(self.dzc)], self.gp["cls"], False, self.conv_impl)
        ops.channel_sum(bv(self.dzc), self.gp["cls.bias"], False)
        ops.conv2d(B, G, G, 1, [bv(self.dzc)], self.wp["cls.dgrad"], None, [bv(self.fb.dout)], [False], None, self.conv_impl)
        # fpn_bottleneck over cat(fpn_0, fpn_1, fpn_2, psp)
        self._cbr_bwd(self.fb, [bv(self.fpn[0].out), bv(self.fpn[1].out), bv(self.fpn[2].out), bv(self.bott.out)], self.fb.dout,
                      [bv(self.fpn[0].dout), bv(self.fpn[1].dout), bv(self.fpn[2].dout), bv(self.bott.dout)], [False] * 4)
        for k in range(3):
            self._cbr_bwd(self.fpn[k], [bv(self.td[k])], self.fpn[k].dout, [bv(self.dtd[k])], [False])
        # top-down sums: td_0 = lat_0 + td_1, td_1 = lat_1 + td_2, td_2 = lat_2 + psp
        ops.bilinear_nhwc_fwd(B, G, G, G, G, self.dtd[0], self.dtd[1], True)
        ops.bilinear_nhwc_fwd(B, G, G, G, G, self.dtd[1], self.dtd[2], True)
        ops.bilinear_nhwc_fwd(B, G, G, G, G, self.dtd[2], self.bott.dout, True)
        # PSP bottleneck over cat(feat_3, up_1, up_2, up_3, up_6): the feature gradient goes to the token-gradient buffer (assign)
        ops.zero_(self.dxn)
        self._cbr_bwd(self.bott, [feats[3]] + [bv(u) for u in self.up], self.bott.dout,
                      [self.grid_view(self.dxn)] + [bv(u) for u in self.dup], [False] * 5)
        for i, s in enumerate(SCALES):
            L = self.psp[i]
            ops.bilinear_nhwc_bwd(B, s, s, G, G, self.dup[i], L.dout, False)
            self._cbr_bwd(L, [L.v(self.pooled[i])], L.dout, [L.v(self.dpooled[i])], [False])
            ops.adaptive_avgpool_bwd(self.dpooled[i], s, self.grid_view(self.dxn), True)
        # lateral modules: weight gradients now, data gradients when the encoder backward reaches their depth (_inject)
        for k in range(3):
            self._cbr_bwd(self.lat[k], [feats[k]], self.dtd[k])

    def _inject(self, bi: int):
        c = bi + 1                              # self.dx is the gradient of the residual stream after block bi
        for k in range(3):
            if self.out_indices[k] == c:
                L = self.lat[k]
                self.ops.conv2d(L.n, L.h, L.w, 1, [L.v(L.dy)], self.wp[f"{L.name}.dgrad"], None, [self.grid_view(self.dx)], [True], None, self.conv_impl)
