"""FinetunerSegmentation with configs["mlp"] = True (models/model_utilities.py:60-65): tokens -> [B, dim, 14, 14] -> bilinear 224 ->
Conv2d(dim, 512, 1) -> ReLU -> Conv2d(512, num_classes, 1).  The first 1x1 conv commutes with the bilinear interpolation (both are
linear, the interpolation weights sum to 1 so the bias commutes too), so it runs on the 14 x 14 grid: the upsampled tensor has 512
channels instead of `dim`, and the dim x 224 x 224 tensor of the reference is never formed.
"""
from __future__ import annotations

import torch

from .lib import View
from .vit_engine import ViTSegEngine

HID, PADK = 512, 32


class ViTMlpHeadEngine(ViTSegEngine):
    def __init__(self, ops, module, enc_prefix, cfg, num_classes, B, H, W, dtype, device):
        super().__init__(ops, module, enc_prefix, cfg, "mlp", num_classes, B, H, W, dtype, device)

    def _alloc(self):
        super()._alloc()
        B, G, D, dev, T = self.B, self.G, self.D, self.device, self.dtype
        HW = self.H * self.W
        z = lambda r, c, dt=None: torch.zeros(r, c, dtype=dt or T, device=dev)
        self.h1, self.dh1 = z(B * G * G, HID), z(B * G * G, HID)           # conv1 output on the token grid (dense [B, G, G, 512])
        self.up, self.dup = z(B * HW, HID), z(B * HW, HID)                 # upsampled (+ ReLU in place) / its gradient
        self.zc, self.dzc = z(B * HW, PADK), z(B * HW, PADK)               # class planes padded to 32 channels (pad stays zero)
        self.wp["mlp0.fwd"], self.wp["mlp0.dgrad"] = z(1, HID * D).view(-1), z(1, D * HID).view(-1)
        self.wp["mlp2.fwd"], self.wp["mlp2.dgrad"] = z(1, PADK * HID).view(-1), z(1, HID * PADK).view(-1)
        self.wp["mlp2.bias"] = torch.zeros(PADK, dtype=torch.float32, device=dev)
        self.gp_mlp2 = torch.zeros(PADK * HID, dtype=torch.float32, device=dev)
        self.gp_mlp2_bias = torch.zeros(PADK, dtype=torch.float32, device=dev)
        self._dz_table = None

    def _pack_jobs(self):
        P, jobs, D, K = self.params, [], self.D, self.K
        for name, (o, i) in self.linears.items():
            w = P.p(f"{name}.weight")
            jobs.append((w, self.wp[f"{name}.fwd"], (o * i,), (1,), 0))
            jobs.append((w, self.wp[f"{name}.dgrad"], (i, o), (1, i), 0))
        w0, w2 = P.p("head.0.weight"), P.p("head.2.weight")
        jobs.append((w0, self.wp["mlp0.fwd"], (HID * D,), (1,), 0))
        jobs.append((w0, self.wp["mlp0.dgrad"], (D, HID), (1, D), 0))
        jobs.append((w2, self.wp["mlp2.fwd"], (K * HID,), (1,), 0))
        jobs.append((w2, self.wp["mlp2.dgrad"], (HID, K), (1, HID), 0, (PADK, 1), 0))
        jobs.append((P.p("head.2.bias"), self.wp["mlp2.bias"], (K,), (1,), 0))
        return jobs

    def _unpack_jobs(self):
        P = self.params
        return [(self.gp_mlp2, P.g("head.2.weight"), (self.K * HID,), (1,), 0),
                (self.gp_mlp2_bias, P.g("head.2.bias"), (self.K,), (1,), 0)]

    def _dense(self, t: torch.Tensor, h: int, w: int) -> View:
        c = t.shape[1]
        return View(t.view(-1), 0, self.B, h, w, c, h * w * c, w * c, c)

    def _head_forward(self, training: bool) -> torch.Tensor:
        ops, P, B, G, H, W = self.ops, self.params, self.B, self.G, self.H, self.W
        ops.conv2d(B, G, G, 1, [self.grid_view(self.tok)], self.wp["mlp0.fwd"], P.p("head.0.bias"), [self._dense(self.h1, G, G)], [False], None,
                   self.conv_impl)
        ops.bilinear_nhwc_fwd(B, G, G, H, W, self.h1, self.up, False)
        ops.relu_fwd(self.up, self.up)
        ops.conv2d(B, H, W, 1, [self._dense(self.up, H, W)], self.wp["mlp2.fwd"], self.wp["mlp2.bias"], [self._dense(self.zc, H, W)], [False], None,
                   self.conv_impl)
        ops.permute_cast(self.zc, self.logits, (B, self.K, H * W), (H * W * PADK, 1, PADK))      # NHWC (padded) -> NCHW fp32
        return self.logits

    def _head_backward(self, dlogits: torch.Tensor):
        ops, P, B, G, H, W, K = self.ops, self.params, self.B, self.G, self.H, self.W, self.K
        HW = H * W
        if self._dz_table is None or self._dz_table[1] != dlogits.data_ptr():
            job = (dlogits.view(-1), self.dzc.view(-1), (B, HW, K), (K * HW, 1, HW), 0, (HW * PADK, PADK, 1), 0)
            self._dz_table = (ops.make_permute_table([job], self.device), dlogits.data_ptr())
        ops.permute_cast_table(self._dz_table[0])                                                 # NCHW fp32 -> NHWC storage dtype
        upv, dzv, dupv = self._dense(self.up, H, W), self._dense(self.dzc, H, W), self._dense(self.dup, H, W)
        ops.conv2d_wgrad(B, H, W, 1, [upv], [dzv], self.gp_mlp2, False, self.conv_impl)
        ops.channel_sum(dzv, self.gp_mlp2_bias, False)
        ops.conv2d(B, H, W, 1, [dzv], self.wp["mlp2.dgrad"], None, [dupv], [False], None, self.conv_impl)
        ops.relu_bwd(self.up, self.dup, self.dup)
        ops.bilinear_nhwc_bwd(B, G, G, H, W, self.dup, self.dh1, False)
        gv, dhv = self.grid_view(self.tok), self._dense(self.dh1, G, G)
        ops.conv2d_wgrad(B, G, G, 1, [gv], [dhv], P.g("head.0.weight"), False, self.conv_impl)
        ops.channel_sum(dhv, P.g("head.0.bias"), False)
        ops.zero_(self.dxn)                                                                        # cls / padding rows get no gradient
        ops.conv2d(B, G, G, 1, [dhv], self.wp["mlp0.dgrad"], None, [self.grid_view(self.dxn)], [False], None, self.conv_impl)


# ---------------------------------------------------------------------------------------------------------------------
# configs["decoder"] = True: head = Decoder (models/model_utilities.py:21-48): ConvTranspose2d(1024, 128, 4, 2, 1) -> ReLU ->
# Upsample(x2, nearest) -> ConvTranspose2d(128, 64, 4, 2, 1) -> ReLU -> ConvTranspose2d(64, K, 4, 2, 1), on the 14 x 14 token map
# (no bilinear interpolation on this branch, :88).  Each ConvTranspose2d(k4, s2, p1) is ONE 3x3 ks_conv2d launch whose output
# channels are the four 2x2 phases (strided destination views), as in the ChangeFormer decoder; the nearest upsample is four
# strided copies of the batched permute kernel (forward) / four accumulating gathers (backward).
# ---------------------------------------------------------------------------------------------------------------------
UP4_TERMS = [(0, 0, 1), (0, -1, 3), (1, 0, 2), (1, 1, 0)]      # (output phase, input offset, kernel index) of ConvTranspose2d(k4, s2, p1)
DEC_CH = (1024, 128, 64)


class _Deconv:
    def __init__(self, name, cin, cout, cpad, hi):
        self.name, self.cin, self.cout, self.cpad, self.hi, self.ho = name, cin, cout, cpad, hi, 2 * hi


class ViTDecoderHeadEngine(ViTSegEngine):
    def __init__(self, ops, module, enc_prefix, cfg, num_classes, B, H, W, dtype, device):
        assert cfg["dim"] == DEC_CH[0], "Decoder.deconv1 is hard-wired to 1024 input channels (model_utilities.py:26)"
        super().__init__(ops, module, enc_prefix, cfg, "decoder", num_classes, B, H, W, dtype, device)

    def _alloc(self):
        super()._alloc()
        B, G, dev, T, K = self.B, self.G, self.device, self.dtype, self.K
        z = lambda r, c: torch.zeros(r, c, dtype=T, device=dev)
        self.dc = [_Deconv("head.deconv1", DEC_CH[0], DEC_CH[1], DEC_CH[1], G), _Deconv("head.deconv2", DEC_CH[1], DEC_CH[2], DEC_CH[2], 4 * G),
                   _Deconv("head.deconv3", DEC_CH[2], K, 16, 8 * G)]
        d1, d2, d3 = self.dc
        self.y1, self.dy1 = z(B * d1.ho ** 2, d1.cpad), z(B * d1.ho ** 2, d1.cpad)          # relu(deconv1): 28 x 28 x 128
        self.u, self.du = z(B * d2.hi ** 2, d1.cpad), z(B * d2.hi ** 2, d1.cpad)             # nearest x2: 56 x 56 x 128
        self.y2, self.dy2 = z(B * d2.ho ** 2, d2.cpad), z(B * d2.ho ** 2, d2.cpad)          # relu(deconv2): 112 x 112 x 64
        self.y3, self.dy3 = z(B * d3.ho ** 2, d3.cpad), z(B * d3.ho ** 2, d3.cpad)          # deconv3: 224 x 224 x (K padded to 16)
        self.dtokg = z(B * G * G, DEC_CH[0])                                                   # d(token map), dense
        self.gpd = {}
        for d in self.dc:
            n4 = 4 * d.cpad
            self.wp[f"{d.name}.fwd"] = torch.zeros(9 * n4 * d.cin, dtype=T, device=dev)
            self.wp[f"{d.name}.dgrad"] = torch.zeros(9 * d.cin * n4, dtype=T, device=dev)
            self.wp[f"{d.name}.bias4"] = torch.zeros(n4, dtype=torch.float32, device=dev)
            self.gpd[d.name] = torch.zeros(9 * n4 * d.cin, dtype=torch.float32, device=dev)
            self.gpd[f"{d.name}.bias"] = torch.zeros(d.cpad, dtype=torch.float32, device=dev)
        self._tabs = None

    def _pack_jobs(self):
        P, jobs = self.params, []
        for name, (o, i) in self.linears.items():
            w = P.p(f"{name}.weight")
            jobs.append((w, self.wp[f"{name}.fwd"], (o * i,), (1,), 0))
            jobs.append((w, self.wp[f"{name}.dgrad"], (i, o), (1, i), 0))
        for d in self.dc:
            w, ci, co, cp = P.p(f"{d.name}.weight"), d.cin, d.cout, d.cpad                   # ConvTranspose2d weight (Cin, Cout, 4, 4)
            for (a, di, ky) in UP4_TERMS:
                for (b_, dj, kx) in UP4_TERMS:
                    ph, off = a * 2 + b_, ky * 4 + kx
                    tap, tapd = (di + 1) * 3 + (dj + 1), (1 - di) * 3 + (1 - dj)
                    jobs.append((w, self.wp[f"{d.name}.fwd"], (co, ci), (16, co * 16), off, (ci, 1), (tap * 4 * cp + ph * cp) * ci))
                    jobs.append((w, self.wp[f"{d.name}.dgrad"], (ci, co), (co * 16, 16), off, (4 * cp, 1), tapd * ci * 4 * cp + ph * cp))
            jobs.append((P.p(f"{d.name}.bias"), self.wp[f"{d.name}.bias4"], (4, co), (0, 1), 0, (cp, 1), 0))
        return jobs

    def _unpack_jobs(self):
        P, jobs = self.params, []
        for d in self.dc:
            ci, co, cp = d.cin, d.cout, d.cpad
            for (a, di, ky) in UP4_TERMS:
                for (b_, dj, kx) in UP4_TERMS:
                    ph, off, tap = a * 2 + b_, ky * 4 + kx, (di + 1) * 3 + (dj + 1)           # grad[ci][co][ky][kx] = gp[tap][ph*cp + co][ci]
                    jobs.append((self.gpd[d.name], P.g(f"{d.name}.weight"), (ci, co), (1, ci), (tap * 4 * cp + ph * cp) * ci, (co * 16, 16), off))
            jobs.append((self.gpd[f"{d.name}.bias"], P.g(f"{d.name}.bias"), (co,), (1,), 0))
        return jobs

    def _dense(self, t: torch.Tensor, h: int) -> View:
        c = t.shape[1]
        return View(t.view(-1), 0, self.B, h, h, c, h * h * c, h * c, c)

    def _phases(self, t: torch.Tensor, ho: int):
        v = self._dense(t, ho)
        return [v.phase(k // 2, k % 2) for k in range(4)]

    def _tables_up(self):
        if self._tabs is None:
            B, h, C = self.B, self.dc[0].ho, DEC_CH[1]
            W2 = 2 * h
            jobs = [(self.y1.view(-1), self.u.view(-1), (B, h, h, C), (h * h * C, h * C, C, 1), 0, (W2 * W2 * C, 2 * W2 * C, 2 * C, 1), (a * W2 + b) * C)
                    for a in (0, 1) for b in (0, 1)]
            K, HW = self.K, self.H * self.W
            self._tabs = self.ops.make_permute_table(jobs, self.device)
        return self._tabs

    def _head_forward(self, training: bool) -> torch.Tensor:
        ops, B, G, K = self.ops, self.B, self.G, self.K
        d1, d2, d3 = self.dc
        ops.conv2d(B, G, G, 3, [self.grid_view(self.tok)], self.wp[f"{d1.name}.fwd"], self.wp[f"{d1.name}.bias4"], self._phases(self.y1, d1.ho), None,
                   None, self.conv_impl)
        ops.relu_fwd(self.y1, self.y1)
        ops.permute_cast_table(self._tables_up())                                               # nearest x2
        ops.conv2d(B, d2.hi, d2.hi, 3, [self._dense(self.u, d2.hi)], self.wp[f"{d2.name}.fwd"], self.wp[f"{d2.name}.bias4"],
                   self._phases(self.y2, d2.ho), None, None, self.conv_impl)
        ops.relu_fwd(self.y2, self.y2)
        ops.conv2d(B, d3.hi, d3.hi, 3, [self._dense(self.y2, d3.hi)], self.wp[f"{d3.name}.fwd"], self.wp[f"{d3.name}.bias4"],
                   self._phases(self.y3, d3.ho), None, None, self.conv_impl)
        HW = self.H * self.W
        ops.permute_cast(self.y3, self.logits, (B, K, HW), (HW * d3.cpad, 1, d3.cpad))         # NHWC (padded) -> NCHW fp32
        return self.logits

    def _deconv_bwd(self, d: _Deconv, src: View, dy: torch.Tensor, dsrc):
        ops = self.ops
        ph = self._phases(dy, d.ho)
        ops.conv2d_wgrad(self.B, d.hi, d.hi, 3, [src], ph, self.gpd[d.name], False, self.conv_impl)
        ops.channel_sum(self._dense(dy, d.ho), self.gpd[f"{d.name}.bias"], False)
        if dsrc is not None:
            ops.conv2d(self.B, d.hi, d.hi, 3, ph, self.wp[f"{d.name}.dgrad"], None, [dsrc], [False], None, self.conv_impl)

    def _head_backward(self, dlogits: torch.Tensor):
        ops, B, G, K = self.ops, self.B, self.G, self.K
        d1, d2, d3 = self.dc
        HW = self.H * self.W
        if getattr(self, "_dz_tab", None) is None or self._dz_tab[1] != dlogits.data_ptr():
            job = (dlogits.view(-1), self.dy3.view(-1), (B, HW, K), (K * HW, 1, HW), 0, (HW * d3.cpad, d3.cpad, 1), 0)
            self._dz_tab = (ops.make_permute_table([job], self.device), dlogits.data_ptr())
        ops.permute_cast_table(self._dz_tab[0])                                                 # NCHW fp32 -> NHWC (pad channels stay zero)
        self._deconv_bwd(d3, self._dense(self.y2, d3.hi), self.dy3, self._dense(self.dy2, d3.hi))
        ops.relu_bwd(self.y2, self.dy2, self.dy2)
        self._deconv_bwd(d2, self._dense(self.u, d2.hi), self.dy2, self._dense(self.du, d2.hi))
        h, C = d1.ho, DEC_CH[1]                                                                  # nearest x2 backward: sum of the four phases
        W2 = 2 * h
        for k, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            ops.permute_cast(self.du, self.dy1, (B, h, h, C), (W2 * W2 * C, 2 * W2 * C, 2 * C, 1), accumulate=(k > 0), src_offset=(a * W2 + b) * C)
        ops.relu_bwd(self.y1, self.dy1, self.dy1)
        self._deconv_bwd(d1, self.grid_view(self.tok), self.dy1, self._dense(self.dtokg, G))
        ops.zero_(self.dxn)
        Dn = DEC_CH[0]
        ops.permute_cast_table(self._tok_table())

    def _tok_table(self):
        if getattr(self, "_tt", None) is None:
            B, G, Dn, Tp = self.B, self.G, DEC_CH[0], self.Tp
            job = (self.dtokg.view(-1), self.dxn.view(-1), (B, G * G, Dn), (G * G * Dn, Dn, 1), 0, (Tp * Dn, Dn, 1), Dn)   # row offset 1: the cls slot
            self._tt = self.ops.make_permute_table([job], self.device)
        return self._tt
