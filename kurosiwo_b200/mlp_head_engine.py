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
