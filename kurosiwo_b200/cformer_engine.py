"""Host-side schedule of the ChangeFormerV6 training step over the C-ABI ops.

Reference path being replaced: models/changeformer.py - ChangeFormerV6.forward :666-676, EncoderTransformer_v3.forward_features
:430-465, Block.forward :244-248, Attention.forward :186-208, Mlp / DWConv :84-133, OverlapPatchEmbed.forward :285-292,
DecoderTransformer_v3.forward :568-641 - and its autograd backward inside training/change_detection_trainer.py:136-177.

Layout: a token matrix [B*N, C] IS the NHWC feature map [B, H, W, C], so the reference's flatten / transpose / reshape / permute
calls (:92-94, :143-145, :192-193, :439) do not exist here.  The encoder has no cross-sample coupling (LayerNorm, per-sample
attention), so both dates run through it as ONE batch of 2B images (first B = date 1); the decoder sees the two halves as views.
  nn.Linear                         -> 1x1 ks_conv2d / ks_conv2d_wgrad over the token matrix (tcgen05 in bf16 mode)
  OverlapPatchEmbed.proj, attn.sr   -> ks_conv2d_strided (+ _dgrad / _wgrad)
  Attention core                    -> ks_xattention_fwd / _bwd (49 reduced keys at 224x224 input)
  DWConv + GELU                     -> ks_dwconv3x3_*, ks_gelu_*
  conv_diff / ResidualBlock / fuse  -> 3x3 / 1x1 ks_conv2d (virtual concat = K-dimension view list), ks_relu_*, ks_bn_* (BatchNorm
                                       AFTER the ReLU, no ReLU after it), ks_bilinear_nhwc_*
  UpsampleConvLayer ConvT(k4,s2,p1) -> ONE 3x3 ks_conv2d launch producing 4*E channels, one group per 2x2 output phase
                                       (out[2i] = w[1] x[i] + w[3] x[i-1], out[2i+1] = w[2] x[i] + w[0] x[i+1])
  ResidualBlock `* 0.1`             -> folded into the packed conv2 weights / bias (permute-table value scale)
  Sigmoid on the outputs            -> ks_sigmoid_head_*; the criterion reads the post-sigmoid map (:635-639, trainer :166-170)
Round-1 restriction: the stochastic regularisers (Dropout 0.1, attention dropout 0.1, DropPath 0.1; :652-654) run with p = 0.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .engine_common import FlatParams, TrainStepMixin
from .lib import IMPL_SIMT, View

EMBED_DIMS, DEPTHS, HEADS, SR = [64, 128, 320, 512], [3, 3, 4, 3], [1, 2, 4, 8], [8, 4, 2, 1]
BN_EPS, BN_MOMENTUM = 1e-5, 0.1
HEAD_PAD = 32          # classifier Cout padded to 32: both its data- and weight-gradient convs then take the tcgen05 path (C % 32 == 0)
# (output phase a, input offset di, kernel index ky) per axis of ConvTranspose2d(k4, s2, p1)
UP4_TERMS = [(0, 0, 1), (0, -1, 3), (1, 0, 2), (1, 1, 0)]


def nhwc(t: torch.Tensor, N: int, H: int, W: int, row0: int = 0) -> View:
    """Rows [row0, row0 + N*H*W) of a [rows, C] matrix as the NHWC view [N, H, W, C]."""
    Cn = t.shape[1]
    return View(t.view(-1), row0 * Cn, N, H, W, Cn, H * W * Cn, W * Cn, Cn)


class _Blk:
    pass


class _Stage:
    pass


class ChangeFormerEngine(TrainStepMixin):
    def __init__(self, ops, module: torch.nn.Module, in_ch: int, num_classes: int, embed_dim: int, decoder_softmax: bool,
                 N: int, H: int, W: int, dtype: torch.dtype, device, conv_impl: int = 0):
        assert num_classes == 3, "the fused head/loss kernels are built for num_classes == 3 (configs/config.json:13)"
        assert H == W and H % 32 == 0, "ChangeFormer needs a square input divisible by 32 (four strided patch embeddings)"
        assert decoder_softmax, "the fused path applies the Sigmoid head (decoder_softmax=true in configs/method/changeformer)"
        self.ops, self.module, self.dtype, self.device = ops, module, dtype, torch.device(device)
        self.in_ch, self.K, self.E, self.N, self.H, self.W = in_ch, num_classes, embed_dim, N, H, W
        self.B2 = 2 * N
        self.conv_impl = conv_impl
        self.params = FlatParams(module)
        self.wp: Dict[str, torch.Tensor] = {}
        self.gp: Dict[str, torch.Tensor] = {}
        self._pack, self._unpack = [], []          # deferred job builders (need the flat parameter buffer)
        self._alloc()

    # ------------------------------------------------------------------------------------------
    def _z(self, rows, C, dtype=None):
        return torch.zeros(rows, C, dtype=dtype or self.dtype, device=self.device)

    def _v(self, rows):
        return torch.zeros(rows, dtype=torch.float32, device=self.device)

    def _buf(self, n, dtype=None):
        return torch.zeros(n, dtype=dtype or self.dtype, device=self.device)

    def _alloc(self):
        B2, N, E, dev = self.B2, self.N, self.E, self.device
        self.x_in = self._z(B2 * self.H * self.W, self.in_ch)              # NHWC copy of cat(x1, x2)
        self.stages: List[_Stage] = []
        hin, cin = self.H, self.in_ch
        for s in range(4):
            st = _Stage()
            st.s, st.C, st.heads, st.sr = s, EMBED_DIMS[s], HEADS[s], SR[s]
            st.Hin, st.Cin = hin, cin
            st.stride = 4 if s == 0 else 2
            st.Hs = (hin + 6 - 7) // st.stride + 1
            st.Ns = st.Hs * st.Hs
            st.R = B2 * st.Ns
            st.dh = st.C // st.heads
            st.Hk = st.Hs // st.sr if st.sr > 1 else st.Hs
            st.Nk = st.Hk * st.Hk
            st.Rk = B2 * st.Nk
            C = st.C
            st.pe = f"Tenc_x2.patch_embed{s + 1}"
            st.pe_y, st.pm, st.pr = self._z(st.R, C), self._v(st.R), self._v(st.R)
            x = self._z(st.R, C)
            st.x0 = x
            st.blocks = []
            for i in range(DEPTHS[s]):
                b = _Blk()
                b.p = f"Tenc_x2.block{s + 1}.{i}"
                b.x = x
                b.xn1, b.m1, b.r1 = self._z(st.R, C), self._v(st.R), self._v(st.R)
                b.q = self._z(st.R, C)
                if st.sr > 1:
                    b.xr, b.mr, b.rr, b.xrn = self._z(st.Rk, C), self._v(st.Rk), self._v(st.Rk), self._z(st.Rk, C)
                else:
                    b.xrn = b.xn1
                b.kv = self._z(st.Rk, 2 * C)
                b.probs = self._buf(B2 * st.heads * st.Ns * st.Nk)
                b.att = self._z(st.R, C)
                b.xm, b.xn2, b.m2, b.r2 = self._z(st.R, C), self._z(st.R, C), self._v(st.R), self._v(st.R)
                b.h1, b.h2, b.h3 = self._z(st.R, 4 * C), self._z(st.R, 4 * C), self._z(st.R, 4 * C)
                x = self._z(st.R, C)
                b.xo = x
                st.blocks.append(b)
            st.xl = x
            st.f, st.mf, st.rf = self._z(st.R, C), self._v(st.R), self._v(st.R)
            # gradient scratch of the stage
            st.df = self._z(st.R, C)
            st.dx, st.dxn, st.dq, st.datt = self._z(st.R, C), self._z(st.R, C), self._z(st.R, C), self._z(st.R, C)
            st.dkv32, st.dkv = self._z(st.Rk, 2 * C, torch.float32), self._z(st.Rk, 2 * C)
            st.dxrn, st.dxr = self._z(st.Rk, C), self._z(st.Rk, C)
            st.dh3, st.dh1 = self._z(st.R, 4 * C), self._z(st.R, 4 * C)
            st.dpe = self._z(st.R, C)
            st.dw9 = torch.zeros(9 * 4 * C, dtype=torch.float32, device=dev)
            self.stages.append(st)
            hin, cin = st.Hs, C
        # ---- decoder ----
        H0 = self.stages[0].Hs
        self.emb, self.demb = {}, {}
        self.d: Dict[int, _Stage] = {}
        for s in range(4):
            st = self.stages[s]
            d = _Stage()
            rows = N * st.Ns
            self.emb[s], self.demb[s] = self._z(2 * rows, E), self._z(2 * rows, E)
            d.y0, d.z, d.y1 = self._z(rows, E), self._z(rows, E), self._z(rows, E)      # y0/y1 hold relu(conv) after the in-place ReLU
            d.x = self._z(rows, E) if s < 3 else d.y1
            d.bn = torch.zeros(4 * E, dtype=torch.float32, device=dev)
            d.up = self._z(N * H0 * H0, E) if s > 0 else d.x
            d.dx, d.dup = self._z(rows, E), (self._z(N * H0 * H0, E) if s > 0 else None)
            d.dy, d.dz = self._z(rows, E), self._z(rows, E)
            # side prediction head (forward only): conv3x3 E->K, ReLU, BN(K), conv3x3 K->K, Sigmoid
            d.p0, d.p1, d.p2 = self._z(rows, self.K), self._z(rows, self.K), self._z(rows, self.K)
            d.pbn = torch.zeros(4 * self.K, dtype=torch.float32, device=dev)
            d.pout = torch.zeros(N, self.K, st.Hs, st.Hs, dtype=torch.float32, device=dev)
            self.d[s] = d
        self.d[0].dup = None
        r0 = N * H0 * H0
        self.yf, self.cf, self.dcf = self._z(r0, E), self._z(r0, E), self._z(r0, E)
        self.fbn = torch.zeros(4 * E, dtype=torch.float32, device=dev)
        self.res = []
        hh = H0
        for tag_up, tag_res in (("convd2x", "dense_2x"), ("convd1x", "dense_1x")):
            r = _Stage()
            r.up, r.res, r.Hi, r.Ho = f"TDec_x2.{tag_up}.conv2d", f"TDec_x2.{tag_res}.0", hh, 2 * hh
            rows = N * r.Ho * r.Ho
            r.c, r.r, r.co = self._z(rows, E), self._z(rows, E), self._z(rows, E)
            r.dco, r.dr = self._z(rows, E), self._z(rows, E)
            self.res.append(r)
            hh *= 2
        rows = N * self.H * self.W
        self.zc, self.dzc = self._z(rows, HEAD_PAD), self._z(rows, HEAD_PAD)
        self.logits = torch.zeros(N, self.K, self.H, self.W, dtype=torch.float32, device=dev)     # post-Sigmoid output
        nbn = 4 * 2 + 1
        self.stats_all = torch.zeros(nbn * 2 * E, dtype=torch.float64, device=dev)
        self.bstats_all = torch.zeros(nbn * 2 * E, dtype=torch.float64, device=dev)
        self.ones, self.zeros = torch.ones(E, dtype=torch.float32, device=dev), torch.zeros(E, dtype=torch.float32, device=dev)
        self.tmp_bias = torch.zeros(4 * E, dtype=torch.float32, device=dev)
        self._declare_weights()

    def _stat(self, idx, C):
        o = idx * 2 * self.E
        return self.stats_all[o:o + 2 * C], self.bstats_all[o:o + 2 * C]

    # ------------------------------------------------------------------------------------------
    # packed weights
    # ------------------------------------------------------------------------------------------
    def _declare_weights(self):
        dev, T, E, K = self.device, self.dtype, self.E, self.K
        zt = lambda n: torch.zeros(n, dtype=T, device=dev)
        zf = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)
        self.linears: Dict[str, tuple] = {}
        for st in self.stages:
            C = st.C
            self.wp[f"{st.pe}.conv"] = zt(49 * C * st.Cin)
            self.gp[f"{st.pe}.conv"] = zf(49 * C * st.Cin)
            for b in st.blocks:
                self.linears[f"{b.p}.attn.q"] = (C, C)
                self.linears[f"{b.p}.attn.kv"] = (2 * C, C)
                self.linears[f"{b.p}.attn.proj"] = (C, C)
                self.linears[f"{b.p}.mlp.fc1"] = (4 * C, C)
                self.linears[f"{b.p}.mlp.fc2"] = (C, 4 * C)
                if st.sr > 1:
                    self.wp[f"{b.p}.attn.sr"] = zt(st.sr * st.sr * C * C)
                    self.gp[f"{b.p}.attn.sr"] = zf(st.sr * st.sr * C * C)
                self.wp[f"{b.p}.dw"] = zf(9 * 4 * C)
        for s in range(4):
            self.linears[f"TDec_x2.linear_c{s + 1}.proj"] = (E, EMBED_DIMS[s])
            for tag, ci in (("0", 2 * E), ("3", E)):
                nm = f"TDec_x2.diff_c{s + 1}.{tag}"
                self.wp[f"{nm}.fwd"], self.wp[f"{nm}.dgrad"], self.gp[nm] = zt(9 * E * ci), zt(9 * E * ci), zf(9 * E * ci)
            self.wp[f"TDec_x2.make_pred_c{s + 1}.0"] = zt(9 * K * E)
            self.wp[f"TDec_x2.make_pred_c{s + 1}.3"] = zt(9 * K * K)
        self.linears["TDec_x2.linear_fuse.0"] = (E, 4 * E)
        for name, (o, i) in self.linears.items():
            self.wp[f"{name}.fwd"], self.wp[f"{name}.dgrad"] = zt(o * i), zt(o * i)
        for r in self.res:
            self.wp[f"{r.up}.fwd"], self.wp[f"{r.up}.dgrad"] = zt(9 * 4 * E * E), zt(9 * E * 4 * E)
            self.wp[f"{r.up}.bias4"], self.gp[r.up] = zf(4 * E), zf(9 * 4 * E * E)
            for tag in ("conv1", "conv2"):
                nm = f"{r.res}.{tag}.conv2d"
                self.wp[f"{nm}.fwd"], self.wp[f"{nm}.dgrad"], self.gp[nm] = zt(9 * E * E), zt(9 * E * E), zf(9 * E * E)
            self.wp[f"{r.res}.conv2.bias01"] = zf(E)
            self.gp[f"{r.res}.conv2.bias"] = zf(E)
        self.wp["cp.fwd"], self.wp["cp.dgrad"], self.wp["cp.bias"] = zt(9 * HEAD_PAD * E), zt(9 * E * HEAD_PAD), zf(HEAD_PAD)
        self.gp["cp"], self.gp["cp.bias"] = zf(9 * HEAD_PAD * E), zf(HEAD_PAD)

    def _pack_jobs(self):
        P, E, K, jobs = self.params, self.E, self.K, []
        for name, (o, i) in self.linears.items():
            w = P.p(f"{name}.weight")
            jobs.append((w, self.wp[f"{name}.fwd"], (o * i,), (1,), 0))
            jobs.append((w, self.wp[f"{name}.dgrad"], (i, o), (1, i), 0))
        for st in self.stages:
            C = st.C
            jobs.append((P.p(f"{st.pe}.proj.weight"), self.wp[f"{st.pe}.conv"], (49, C, st.Cin), (1, st.Cin * 49, 49), 0))
            for b in st.blocks:
                if st.sr > 1:
                    kk = st.sr * st.sr
                    jobs.append((P.p(f"{b.p}.attn.sr.weight"), self.wp[f"{b.p}.attn.sr"], (kk, C, C), (1, C * kk, kk), 0))
                jobs.append((P.p(f"{b.p}.mlp.dwconv.dwconv.weight"), self.wp[f"{b.p}.dw"], (9, 4 * C), (1, 9), 0))
        def conv3(nm, wname, co, ci, scale=None):
            w = P.p(wname)
            jobs.append((w, self.wp[f"{nm}.fwd"], (9, co, ci), (1, ci * 9, 9), 0, None, 0, scale))
            jobs.append((w, self.wp[f"{nm}.dgrad"], (9, ci, co), (-1, 9, ci * 9), 8, None, 0, scale))
        for s in range(4):
            conv3(f"TDec_x2.diff_c{s + 1}.0", f"TDec_x2.diff_c{s + 1}.0.weight", E, 2 * E)
            conv3(f"TDec_x2.diff_c{s + 1}.3", f"TDec_x2.diff_c{s + 1}.3.weight", E, E)
            jobs.append((P.p(f"TDec_x2.make_pred_c{s + 1}.0.weight"), self.wp[f"TDec_x2.make_pred_c{s + 1}.0"], (9, K, E), (1, E * 9, 9), 0))
            jobs.append((P.p(f"TDec_x2.make_pred_c{s + 1}.3.weight"), self.wp[f"TDec_x2.make_pred_c{s + 1}.3"], (9, K, K), (1, K * 9, 9), 0))
        for r in self.res:
            w = P.p(f"{r.up}.weight")                       # ConvTranspose2d (Cin, Cout, 4, 4)
            for (a, di, ky) in UP4_TERMS:
                for (b_, dj, kx) in UP4_TERMS:
                    ph, off = a * 2 + b_, ky * 4 + kx
                    tap = (di + 1) * 3 + (dj + 1)
                    jobs.append((w, self.wp[f"{r.up}.fwd"], (E, E), (16, E * 16), off, (E, 1), (tap * 4 * E + ph * E) * E))
                    tapd = (1 - di) * 3 + (1 - dj)
                    jobs.append((w, self.wp[f"{r.up}.dgrad"], (E, E), (E * 16, 16), off, (4 * E, 1), tapd * E * 4 * E + ph * E))
            jobs.append((P.p(f"{r.up}.bias"), self.wp[f"{r.up}.bias4"], (4, E), (0, 1), 0))
            conv3(f"{r.res}.conv1.conv2d", f"{r.res}.conv1.conv2d.weight", E, E)
            conv3(f"{r.res}.conv2.conv2d", f"{r.res}.conv2.conv2d.weight", E, E, 0.1)            # `out = conv2(out) * 0.1` (:481)
            jobs.append((P.p(f"{r.res}.conv2.conv2d.bias"), self.wp[f"{r.res}.conv2.bias01"], (E,), (1,), 0, None, 0, 0.1))
        w = P.p("TDec_x2.change_probability.conv2d.weight")  # (K, E, 3, 3), output channels padded to HEAD_PAD
        jobs.append((w, self.wp["cp.fwd"], (9, K, E), (1, E * 9, 9), 0, (HEAD_PAD * E, E, 1), 0))
        jobs.append((w, self.wp["cp.dgrad"], (9, E, K), (-1, 9, E * 9), 8, (E * HEAD_PAD, HEAD_PAD, 1), 0))
        jobs.append((P.p("TDec_x2.change_probability.conv2d.bias"), self.wp["cp.bias"], (K,), (1,), 0))
        return jobs

    def _unpack_jobs(self):
        P, E, K, jobs = self.params, self.E, self.K, []
        for st in self.stages:
            C = st.C
            jobs.append((self.gp[f"{st.pe}.conv"], P.g(f"{st.pe}.proj.weight"), (C, st.Cin, 49), (st.Cin, 1, C * st.Cin), 0))
            for b in st.blocks:
                if st.sr > 1:
                    kk = st.sr * st.sr
                    jobs.append((self.gp[f"{b.p}.attn.sr"], P.g(f"{b.p}.attn.sr.weight"), (C, C, kk), (C, 1, C * C), 0))
        def conv3(nm, wname, co, ci, scale=None):
            jobs.append((self.gp[nm], P.g(wname), (co, ci, 9), (ci, 1, co * ci), 0, None, 0, scale))
        for s in range(4):
            conv3(f"TDec_x2.diff_c{s + 1}.0", f"TDec_x2.diff_c{s + 1}.0.weight", E, 2 * E)
            conv3(f"TDec_x2.diff_c{s + 1}.3", f"TDec_x2.diff_c{s + 1}.3.weight", E, E)
        for r in self.res:
            for (a, di, ky) in UP4_TERMS:
                for (b_, dj, kx) in UP4_TERMS:
                    ph, off = a * 2 + b_, ky * 4 + kx
                    tap = (di + 1) * 3 + (dj + 1)          # grad[ci][co][ky][kx] = gp[tap][ph*E+co][ci]
                    jobs.append((self.gp[r.up], P.g(f"{r.up}.weight"), (E, E), (1, E), (tap * 4 * E + ph * E) * E, (E * 16, 16), off))
            conv3(f"{r.res}.conv1.conv2d", f"{r.res}.conv1.conv2d.weight", E, E)
            conv3(f"{r.res}.conv2.conv2d", f"{r.res}.conv2.conv2d.weight", E, E, 0.1)
            jobs.append((self.gp[f"{r.res}.conv2.bias"], P.g(f"{r.res}.conv2.conv2d.bias"), (E,), (1,), 0, None, 0, 0.1))
        jobs.append((self.gp["cp"], P.g("TDec_x2.change_probability.conv2d.weight"), (K, E, 9), (E, 1, HEAD_PAD * E), 0))
        jobs.append((self.gp["cp.bias"], P.g("TDec_x2.change_probability.conv2d.bias"), (K,), (1,), 0))
        return jobs

    def _tables(self):
        key = (self.params.flat.data_ptr(), self.params.grad.data_ptr())
        if getattr(self, "_table_key", None) != key:
            self._pack_table = self.ops.make_permute_table(self._pack_jobs(), self.device)
            self._unpack_table = self.ops.make_permute_table(self._unpack_jobs(), self.device)
            self._table_key = key
        return self._pack_table, self._unpack_table

    # ------------------------------------------------------------------------------------------
    # small helpers
    # ------------------------------------------------------------------------------------------
    def _lv(self, t: torch.Tensor) -> View:
        """A [rows, C] matrix as the view the 1x1 conv engine tiles best: [1, rows/16, 16, C] when rows % 16 == 0."""
        R, Cn = t.shape
        w = next(w for w in (16, 8, 4, 2, 1) if R % w == 0)
        return View(t.view(-1), 0, 1, R // w, w, Cn, R * Cn, w * Cn, Cn)

    def _lin(self, a, name, out, bias=True, acc=False):
        va, vo = self._lv(a), self._lv(out)
        self.ops.conv2d(va.N, va.H, va.W, 1, [va], self.wp[f"{name}.fwd"], self.params.p(f"{name}.bias") if bias else None, [vo], [acc], None,
                        self.conv_impl)

    def _lin_bwd(self, a, name, dy, da, acc_da=False, bias=True):
        P, va, vy = self.params, self._lv(a), self._lv(dy)
        self.ops.conv2d_wgrad(va.N, va.H, va.W, 1, [va], [vy], P.g(f"{name}.weight"), False, self.conv_impl)
        if bias:
            self._colsum(dy, P.g(f"{name}.bias"))
        if da is not None:
            vd = self._lv(da)
            self.ops.conv2d(vy.N, vy.H, vy.W, 1, [vy], self.wp[f"{name}.dgrad"], None, [vd], [acc_da], None, self.conv_impl)

    def _colsum(self, m: torch.Tensor, out: torch.Tensor):
        v = self._lv(m)
        for c0 in range(0, v.C, 1024):
            c = min(1024, v.C - c0)
            self.ops.channel_sum(v.ch(c0, c), out[c0:c0 + c], False)

    def _bn_eval(self, name, bn):
        C = bn.numel() // 4
        m = self.module.get_submodule(name)
        sc, sh = bn[:C], bn[C:2 * C]
        torch.mul(self.params.p(f"{name}.weight"), torch.rsqrt(m.running_var + BN_EPS), out=sc)
        torch.sub(self.params.p(f"{name}.bias"), m.running_mean * sc, out=sh)

    def _bn_train(self, name, bn, stats, count):
        C = bn.numel() // 4
        m = self.module.get_submodule(name)
        sc, sh, mu, rs = [bn[i * C:(i + 1) * C] for i in range(4)]
        self.ops.bn_finalize(C, float(count), stats, self.params.p(f"{name}.weight"), self.params.p(f"{name}.bias"), BN_EPS, BN_MOMENTUM,
                             m.running_mean, m.running_var, sc, sh, mu, rs)

    def _ensure_nbt(self):
        names = [f"TDec_x2.diff_c{s}.2" for s in (4, 3, 2, 1)] + [f"TDec_x2.make_pred_c{s}.2" for s in (4, 3, 2, 1)] + ["TDec_x2.linear_fuse.1"]
        mods = [self.module.get_submodule(n) for n in names]
        flat = getattr(self, "nbt_all", None)
        if flat is not None and flat.device == self.device and all(m.num_batches_tracked.data_ptr() == flat.data_ptr() + 8 * i for i, m in enumerate(mods)):
            return
        flat = torch.zeros(len(mods), dtype=torch.int64, device=self.device)
        for i, m in enumerate(mods):
            flat[i] = m.num_batches_tracked.to(self.device)
            m._buffers["num_batches_tracked"] = flat[i]
        self.nbt_all = flat

    # ------------------------------------------------------------------------------------------
    def forward(self, x1: torch.Tensor, x2: torch.Tensor, training: bool = True) -> torch.Tensor:
        ops, P, N, B2, E, K = self.ops, self.params, self.N, self.B2, self.E, self.K
        assert tuple(x1.shape) == (N, self.in_ch, self.H, self.W) and tuple(x2.shape) == tuple(x1.shape), \
            f"engine was planned for {(N, self.in_ch, self.H, self.W)}, got {tuple(x1.shape)}"
        P.ensure(self.device)
        Cin, H, W = self.in_ch, self.H, self.W
        for k, x in enumerate((x1, x2)):
            x = x.contiguous()
            if x.dtype != torch.float32:
                x = x.float()
            ops.permute_cast(x, self.x_in.view(-1)[k * N * H * W * Cin:], (N, H, W, Cin), (Cin * H * W, W, 1, H * W))
        ops.permute_cast_table(self._tables()[0])
        if training:
            ops.zero_(self.stats_all)
            self._ensure_nbt()
            self.nbt_all.add_(1)
        # ---------------- encoder (both dates as one batch) ----------------
        src = nhwc(self.x_in, B2, H, W)
        for st in self.stages:
            C = st.C
            ops.conv2d_strided(B2, st.Hin, st.Hin, st.Hs, st.Hs, 7, st.stride, 3, src, self.wp[f"{st.pe}.conv"], P.p(f"{st.pe}.proj.bias"),
                               nhwc(st.pe_y, B2, st.Hs, st.Hs))
            ops.layernorm_fwd(st.pe_y, P.p(f"{st.pe}.norm.weight"), P.p(f"{st.pe}.norm.bias"), 1e-5, st.x0, st.pm, st.pr, None)
            for b in st.blocks:
                ops.layernorm_fwd(b.x, P.p(f"{b.p}.norm1.weight"), P.p(f"{b.p}.norm1.bias"), 1e-6, b.xn1, b.m1, b.r1, b.xm)
                self._lin(b.xn1, f"{b.p}.attn.q", b.q)
                if st.sr > 1:
                    ops.conv2d_strided(B2, st.Hs, st.Hs, st.Hk, st.Hk, st.sr, st.sr, 0, nhwc(b.xn1, B2, st.Hs, st.Hs), self.wp[f"{b.p}.attn.sr"],
                                       P.p(f"{b.p}.attn.sr.bias"), nhwc(b.xr, B2, st.Hk, st.Hk))
                    ops.layernorm_fwd(b.xr, P.p(f"{b.p}.attn.norm.weight"), P.p(f"{b.p}.attn.norm.bias"), 1e-5, b.xrn, b.mr, b.rr, None)
                self._lin(b.xrn, f"{b.p}.attn.kv", b.kv)
                ops.xattention_fwd(B2, st.Ns, st.Nk, st.heads, st.dh, b.q, b.kv, st.dh ** -0.5, b.att, b.probs)
                self._lin(b.att, f"{b.p}.attn.proj", b.xm, acc=True)                              # x = x + attn(norm1(x))
                ops.layernorm_fwd(b.xm, P.p(f"{b.p}.norm2.weight"), P.p(f"{b.p}.norm2.bias"), 1e-6, b.xn2, b.m2, b.r2, b.xo)
                self._lin(b.xn2, f"{b.p}.mlp.fc1", b.h1)
                ops.dwconv3x3_fwd(B2, st.Hs, st.Hs, b.h1, self.wp[f"{b.p}.dw"], P.p(f"{b.p}.mlp.dwconv.dwconv.bias"), b.h2)
                ops.gelu_fwd(b.h2, b.h3)
                self._lin(b.h3, f"{b.p}.mlp.fc2", b.xo, acc=True)                                  # x = x + mlp(norm2(x))
            ops.layernorm_fwd(st.xl, P.p(f"Tenc_x2.norm{st.s + 1}.weight"), P.p(f"Tenc_x2.norm{st.s + 1}.bias"), 1e-6, st.f, st.mf, st.rf, None)
            src = nhwc(st.f, B2, st.Hs, st.Hs)
        # ---------------- decoder ----------------
        H0 = self.stages[0].Hs
        for s in (3, 2, 1, 0):
            st, d = self.stages[s], self.d[s]
            Hs, rows = st.Hs, N * st.Ns
            nm = f"TDec_x2.diff_c{s + 1}"
            self._lin(st.f, f"TDec_x2.linear_c{s + 1}.proj", self.emb[s])
            e1, e2 = nhwc(self.emb[s], N, Hs, Hs, 0), nhwc(self.emb[s], N, Hs, Hs, rows)
            ops.conv2d(N, Hs, Hs, 3, [e1, e2], self.wp[f"{nm}.0.fwd"], P.p(f"{nm}.0.bias"), [nhwc(d.y0, N, Hs, Hs)], None, None, self.conv_impl)
            ops.relu_fwd(d.y0, d.y0)
            if training:
                stats, _ = self._stat(2 * s, E)
                ops.bn_stats(nhwc(d.y0, N, Hs, Hs), stats)
                self._bn_train(f"{nm}.2", d.bn, stats, rows)
            else:
                self._bn_eval(f"{nm}.2", d.bn)
            ops.bn_act(nhwc(d.y0, N, Hs, Hs), d.bn[:E], d.bn[E:2 * E], None, False, nhwc(d.z, N, Hs, Hs), None)
            ops.conv2d(N, Hs, Hs, 3, [nhwc(d.z, N, Hs, Hs)], self.wp[f"{nm}.3.fwd"], P.p(f"{nm}.3.bias"), [nhwc(d.y1, N, Hs, Hs)], None, None,
                       self.conv_impl)
            ops.relu_fwd(d.y1, d.y1)
            if s < 3:                                                    # + F.interpolate(_c{s+1}, scale_factor=2, 'bilinear')
                d.x.copy_(d.y1)
                Hc = self.stages[s + 1].Hs
                ops.bilinear_nhwc_fwd(N, Hc, Hc, Hs, Hs, self.d[s + 1].x, d.x, True)
            # side prediction (no gradient without multi_scale_train; kept for the 5-output contract and the BN state)
            pn = f"TDec_x2.make_pred_c{s + 1}"
            ops.conv2d(N, Hs, Hs, 3, [nhwc(d.x, N, Hs, Hs)], self.wp[f"{pn}.0"], P.p(f"{pn}.0.bias"), [nhwc(d.p0, N, Hs, Hs)], None, None, IMPL_SIMT)
            d.p0.clamp_(min=0)
            if training:
                stats, _ = self._stat(2 * s + 1, K)
                ops.bn_stats(nhwc(d.p0, N, Hs, Hs), stats)
                self._bn_train(f"{pn}.2", d.pbn, stats, rows)
            else:
                self._bn_eval(f"{pn}.2", d.pbn)
            ops.bn_act(nhwc(d.p0, N, Hs, Hs), d.pbn[:K], d.pbn[K:2 * K], None, False, nhwc(d.p1, N, Hs, Hs), None)
            ops.conv2d(N, Hs, Hs, 3, [nhwc(d.p1, N, Hs, Hs)], self.wp[f"{pn}.3"], P.p(f"{pn}.3.bias"), [nhwc(d.p2, N, Hs, Hs)], None, None, IMPL_SIMT)
            ops.sigmoid_head_fwd(nhwc(d.p2, N, Hs, Hs), K, d.pout)
            if s > 0:
                ops.bilinear_nhwc_fwd(N, Hs, Hs, H0, H0, d.x, d.up, False)
        cat = [nhwc(self.d[s].up, N, H0, H0) for s in (3, 2, 1, 0)]       # (_c4_up, _c3_up, _c2_up, _c1) (:613)
        fstats, _ = self._stat(8, E)
        ops.conv2d(N, H0, H0, 1, cat, self.wp["TDec_x2.linear_fuse.0.fwd"], P.p("TDec_x2.linear_fuse.0.bias"), [nhwc(self.yf, N, H0, H0)], None,
                   fstats if training else None, self.conv_impl)
        if training:
            self._bn_train("TDec_x2.linear_fuse.1", self.fbn, fstats, N * H0 * H0)
        else:
            self._bn_eval("TDec_x2.linear_fuse.1", self.fbn)
        ops.bn_act(nhwc(self.yf, N, H0, H0), self.fbn[:E], self.fbn[E:2 * E], None, False, nhwc(self.cf, N, H0, H0), None)
        cur = self.cf
        for r in self.res:
            vin, vc = nhwc(cur, N, r.Hi, r.Hi), nhwc(r.c, N, r.Ho, r.Ho)
            ops.conv2d(N, r.Hi, r.Hi, 3, [vin], self.wp[f"{r.up}.fwd"], self.wp[f"{r.up}.bias4"], [vc.phase(k // 2, k % 2) for k in range(4)], None, None,
                       self.conv_impl)
            ops.conv2d(N, r.Ho, r.Ho, 3, [vc], self.wp[f"{r.res}.conv1.conv2d.fwd"], P.p(f"{r.res}.conv1.conv2d.bias"), [nhwc(r.r, N, r.Ho, r.Ho)], None,
                       None, self.conv_impl)
            ops.relu_fwd(r.r, r.r)
            r.co.copy_(r.c)                                               # residual
            ops.conv2d(N, r.Ho, r.Ho, 3, [nhwc(r.r, N, r.Ho, r.Ho)], self.wp[f"{r.res}.conv2.conv2d.fwd"], self.wp[f"{r.res}.conv2.bias01"],
                       [nhwc(r.co, N, r.Ho, r.Ho)], [True], None, self.conv_impl)
            cur = r.co
        ops.conv2d(N, H, W, 3, [nhwc(cur, N, H, W)], self.wp["cp.fwd"], self.wp["cp.bias"], [nhwc(self.zc, N, H, W)], None, None, self.conv_impl)
        ops.sigmoid_head_fwd(nhwc(self.zc, N, H, W), K, self.logits)
        return self.logits

    def outputs(self) -> List[torch.Tensor]:
        """The reference's 5-element output list [p_c4, p_c3, p_c2, p_c1, cp] (:586-633) of the last forward."""
        return [self.d[s].pout for s in (3, 2, 1, 0)] + [self.logits]

    # ------------------------------------------------------------------------------------------
    def _conv3_bwd(self, nm, wname, srcs: List[View], dy: View, gdsts: Optional[List[View]], gacc=None, bias_grad: Optional[torch.Tensor] = None):
        ops, N_, H_, W_ = self.ops, dy.N, dy.H, dy.W
        ops.conv2d_wgrad(N_, H_, W_, 3, srcs, [dy], self.gp[nm], False, self.conv_impl)
        if bias_grad is not None:
            ops.channel_sum(dy, bias_grad, False)
        if gdsts is not None:
            ops.conv2d(N_, H_, W_, 3, [dy], self.wp[f"{nm}.dgrad"], None, gdsts, gacc, None, self.conv_impl)

    def _bn_bwd_norelu(self, name, bn, dout: View, y: View, bstats, count, dy: View):
        """BatchNorm without a following ReLU: the reduce pass gets (scale, shift) = (0, 1) so that its mask is always on."""
        C, P = bn.numel() // 4, self.params
        mu, rs = bn[2 * C:3 * C], bn[3 * C:4 * C]
        self.ops.bn_bwd_reduce(dout, None, y, self.zeros[:C], self.ones[:C], mu, rs, bstats)
        self.ops.bn_bwd_apply(dout, True, y, None, None, mu, rs, P.p(f"{name}.weight"), bstats, float(count), None, dy,
                              P.g(f"{name}.weight"), P.g(f"{name}.bias"), None, False)

    def backward(self, dout: torch.Tensor):
        ops, P, N, B2, E, K, H, W = self.ops, self.params, self.N, self.B2, self.E, self.K, self.H, self.W
        ops.zero_(P.grad)
        ops.zero_(self.bstats_all)
        H0 = self.stages[0].Hs
        # ---- head + residual / transposed-conv stack ----
        ops.sigmoid_head_bwd(self.logits, dout, K, nhwc(self.dzc, N, H, W))
        last = self.res[-1]
        ops.conv2d_wgrad(N, H, W, 3, [nhwc(last.co, N, H, W)], [nhwc(self.dzc, N, H, W)], self.gp["cp"], False, self.conv_impl)
        ops.channel_sum(nhwc(self.dzc, N, H, W), self.gp["cp.bias"], False)
        ops.conv2d(N, H, W, 3, [nhwc(self.dzc, N, H, W)], self.wp["cp.dgrad"], None, [nhwc(last.dco, N, H, W)], [False], None, self.conv_impl)
        for ri in (1, 0):
            r = self.res[ri]
            vdco, vdr = nhwc(r.dco, N, r.Ho, r.Ho), nhwc(r.dr, N, r.Ho, r.Ho)
            # out = c + 0.1*(conv2(relu(conv1(c))) + b2)
            self._conv3_bwd(f"{r.res}.conv2.conv2d", None, [nhwc(r.r, N, r.Ho, r.Ho)], vdco, [vdr], [False], self.gp[f"{r.res}.conv2.bias"])
            ops.relu_bwd(r.r, r.dr, r.dr)
            self._conv3_bwd(f"{r.res}.conv1.conv2d", None, [nhwc(r.c, N, r.Ho, r.Ho)], vdr, [vdco], [True], P.g(f"{r.res}.conv1.conv2d.bias"))
            # ConvTranspose2d(k4,s2,p1) backward: r.dco is now d(c)
            phases = [vdco.phase(k // 2, k % 2) for k in range(4)]
            src = self.cf if ri == 0 else self.res[0].co
            dsrc = self.dcf if ri == 0 else self.res[0].dco
            ops.conv2d(N, r.Hi, r.Hi, 3, phases, self.wp[f"{r.up}.dgrad"], None, [nhwc(dsrc, N, r.Hi, r.Hi)], [False], None, self.conv_impl)
            ops.conv2d_wgrad(N, r.Hi, r.Hi, 3, [nhwc(src, N, r.Hi, r.Hi)], phases, self.gp[r.up], False, self.conv_impl)
            ops.channel_sum(vdco, P.g(f"{r.up}.bias"), False)
        # ---- linear_fuse: 1x1 conv + BN ----
        _, fb = self._stat(8, E)
        vyf = nhwc(self.yf, N, H0, H0)
        self._bn_bwd_norelu("TDec_x2.linear_fuse.1", self.fbn, nhwc(self.dcf, N, H0, H0), vyf, fb, N * H0 * H0, vyf)     # d(yf) over yf
        cat = [nhwc(self.d[s].up, N, H0, H0) for s in (3, 2, 1, 0)]
        ops.conv2d_wgrad(N, H0, H0, 1, cat, [vyf], P.g("TDec_x2.linear_fuse.0.weight"), False, self.conv_impl)
        gd = [nhwc(self.d[s].dup, N, H0, H0) for s in (3, 2, 1)] + [nhwc(self.d[0].dx, N, H0, H0)]
        ops.conv2d(N, H0, H0, 1, [vyf], self.wp["TDec_x2.linear_fuse.0.dgrad"], None, gd, [False] * 4, None, self.conv_impl)
        # ---- per-scale difference modules, fine -> coarse ----
        for s in (0, 1, 2, 3):
            st, d = self.stages[s], self.d[s]
            Hs, rows = st.Hs, N * st.Ns
            nm = f"TDec_x2.diff_c{s + 1}"
            if s < 3:                      # x_s = relu(y1) + up2(x_{s+1});  up_{s+1} = resize(x_{s+1})
                Hc = self.stages[s + 1].Hs
                ops.bilinear_nhwc_bwd(N, Hc, Hc, Hs, Hs, d.dx, self.d[s + 1].dx, False)
                ops.bilinear_nhwc_bwd(N, Hc, Hc, H0, H0, self.d[s + 1].dup, self.d[s + 1].dx, True)
            ops.relu_bwd(d.y1, d.dx, d.dy)
            vdy, vdz = nhwc(d.dy, N, Hs, Hs), nhwc(d.dz, N, Hs, Hs)
            self._conv3_bwd(f"{nm}.3", None, [nhwc(d.z, N, Hs, Hs)], vdy, [vdz], [False], P.g(f"{nm}.3.bias"))
            _, bs = self._stat(2 * s, E)
            self._bn_bwd_norelu(f"{nm}.2", d.bn, vdz, nhwc(d.y0, N, Hs, Hs), bs, rows, vdy)                 # d(relu(y0)) -> d.dy
            ops.relu_bwd(d.y0, d.dy, d.dy)
            e1, e2 = nhwc(self.emb[s], N, Hs, Hs, 0), nhwc(self.emb[s], N, Hs, Hs, rows)
            g1, g2 = nhwc(self.demb[s], N, Hs, Hs, 0), nhwc(self.demb[s], N, Hs, Hs, rows)
            self._conv3_bwd(f"{nm}.0", None, [e1, e2], vdy, [g1, g2], [False, False], P.g(f"{nm}.0.bias"))
            self._lin_bwd(st.f, f"TDec_x2.linear_c{s + 1}.proj", self.demb[s], st.df)
        # ---- encoder, coarse -> fine ----
        for st in reversed(self.stages):
            C = st.C
            ops.layernorm_bwd(st.df, st.xl, st.mf, st.rf, P.p(f"Tenc_x2.norm{st.s + 1}.weight"), st.dx, False,
                              P.g(f"Tenc_x2.norm{st.s + 1}.weight"), P.g(f"Tenc_x2.norm{st.s + 1}.bias"))
            for b in reversed(st.blocks):
                self._lin_bwd(b.h3, f"{b.p}.mlp.fc2", st.dx, st.dh3)
                ops.gelu_bwd(b.h2, st.dh3, st.dh3)
                ops.zero_(st.dw9)
                ops.dwconv3x3_bwd(B2, st.Hs, st.Hs, b.h1, st.dh3, self.wp[f"{b.p}.dw"], st.dh1, st.dw9, P.g(f"{b.p}.mlp.dwconv.dwconv.bias"))
                ops.permute_cast(st.dw9, P.g(f"{b.p}.mlp.dwconv.dwconv.weight"), (4 * C, 9), (1, 4 * C))       # [9][4C] -> (4C,1,3,3)
                self._lin_bwd(b.xn2, f"{b.p}.mlp.fc1", st.dh1, st.dxn)
                ops.layernorm_bwd(st.dxn, b.xm, b.m2, b.r2, P.p(f"{b.p}.norm2.weight"), st.dx, True, P.g(f"{b.p}.norm2.weight"), P.g(f"{b.p}.norm2.bias"))
                self._lin_bwd(b.att, f"{b.p}.attn.proj", st.dx, st.datt)
                ops.xattention_bwd(B2, st.Ns, st.Nk, st.heads, st.dh, b.q, b.kv, b.probs, st.datt, st.dh ** -0.5, st.dq, st.dkv32)
                ops.permute_cast(st.dkv32, st.dkv, (st.Rk * 2 * C,), (1,))
                if st.sr > 1:
                    self._lin_bwd(b.xrn, f"{b.p}.attn.kv", st.dkv, st.dxrn)
                    ops.layernorm_bwd(st.dxrn, b.xr, b.mr, b.rr, P.p(f"{b.p}.attn.norm.weight"), st.dxr, False,
                                      P.g(f"{b.p}.attn.norm.weight"), P.g(f"{b.p}.attn.norm.bias"))
                    vx, vdr = nhwc(b.xn1, B2, st.Hs, st.Hs), nhwc(st.dxr, B2, st.Hk, st.Hk)
                    ops.conv2d_strided_wgrad(B2, st.Hs, st.Hs, st.Hk, st.Hk, st.sr, st.sr, 0, vx, vdr, self.gp[f"{b.p}.attn.sr"], False)
                    self._colsum(st.dxr, P.g(f"{b.p}.attn.sr.bias"))
                    ops.conv2d_strided_dgrad(B2, st.Hs, st.Hs, st.Hk, st.Hk, st.sr, st.sr, 0, vdr, self.wp[f"{b.p}.attn.sr"], nhwc(st.dxn, B2, st.Hs, st.Hs), False)
                else:
                    self._lin_bwd(b.xrn, f"{b.p}.attn.kv", st.dkv, st.dxn)
                self._lin_bwd(b.xn1, f"{b.p}.attn.q", st.dq, st.dxn, acc_da=True)
                ops.layernorm_bwd(st.dxn, b.x, b.m1, b.r1, P.p(f"{b.p}.norm1.weight"), st.dx, True, P.g(f"{b.p}.norm1.weight"), P.g(f"{b.p}.norm1.bias"))
            ops.layernorm_bwd(st.dx, st.pe_y, st.pm, st.pr, P.p(f"{st.pe}.norm.weight"), st.dpe, False, P.g(f"{st.pe}.norm.weight"), P.g(f"{st.pe}.norm.bias"))
            vsrc = nhwc(self.x_in, B2, H, W) if st.s == 0 else nhwc(self.stages[st.s - 1].f, B2, st.Hin, st.Hin)
            vdpe = nhwc(st.dpe, B2, st.Hs, st.Hs)
            ops.conv2d_strided_wgrad(B2, st.Hin, st.Hin, st.Hs, st.Hs, 7, st.stride, 3, vsrc, vdpe, self.gp[f"{st.pe}.conv"], False)
            self._colsum(st.dpe, P.g(f"{st.pe}.proj.bias"))
            if st.s > 0:
                prev = self.stages[st.s - 1]
                ops.conv2d_strided_dgrad(B2, st.Hin, st.Hin, st.Hs, st.Hs, 7, st.stride, 3, vdpe, self.wp[f"{st.pe}.conv"],
                                         nhwc(prev.df, B2, prev.Hs, prev.Hs), True)
        ops.permute_cast_table(self._tables()[1])
