"""Host-side schedule of the FloodViT training step (ViT encoder + segmentation head) over the C-ABI ops.

Reference path being replaced: models/vision_transformer.py:139-156 (ViT.forward), :84-89 (Transformer.forward), :53-66
(Attention.forward), :19-32 (FeedForward) and models/model_utilities.py:80-94 (FinetunerSegmentation.forward) plus their
autograd backward, inside training/segmentation_trainer.py:54-164.

Layout in HBM: every token matrix is [R = B*Tp, C] row-major in the storage dtype (bf16 perf / fp32 parity), Tp = tokens per
image padded to a multiple of 16 (197 -> 208), cls token first as in the reference (`torch.cat((cls_tokens, x), dim=1)`).
Seen as an NHWC tensor [1, R/16, 16, C] such a matrix is exactly what the tcgen05 conv engine tiles (8 x 16 pixel windows), so
every nn.Linear is a 1x1 ks_conv2d (bias and the residual += in its epilogue), its weight gradient a 1x1 ks_conv2d_wgrad that
lands directly in the flat fp32 gradient buffer ([out][in] is nn.Linear's own layout), its data gradient a ks_conv2d with the
transposed weight.  Padding rows carry finite junk forward and exact zeros backward (attention masks keys >= T).
The residual stream is NOT updated in place: each LayerNorm pass also copies x into the next stream buffer, which the following
projection accumulates into - so every LayerNorm input is still there for the backward.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .engine_common import FlatParams, TrainStepMixin
from .lib import View

LN_EPS = 1e-5
HEAD_PAD = 16


def tok_view(t: torch.Tensor) -> View:
    """[R, C] row-major matrix as the NHWC view [1, R/16, 16, C]."""
    R, Cn = t.shape
    assert R % 16 == 0 and t.is_contiguous()
    return View(t.view(-1), 0, 1, R // 16, 16, Cn, R * Cn, 16 * Cn, Cn)


class _Block:
    __slots__ = ("xa", "xn1", "m1", "r1", "qkv", "probs", "att", "xm", "xn2", "m2", "r2", "u", "h", "pa", "pf")


class ViTSegEngine(TrainStepMixin):
    def __init__(self, ops, module: torch.nn.Module, enc_prefix: str, cfg: dict, head: str, num_classes: int,
                 B: int, H: int, W: int, dtype: torch.dtype, device, conv_impl: int = 0):
        assert num_classes == 3, "the fused head/loss kernels are built for num_classes == 3 (configs/config.json:13)"
        assert head in ("linear", "mlp", "decoder", "upernet"), "heads on the fused path: linear / mlp / decoder (FinetunerSegmentation), upernet (HF UperNetHead)"
        self.head_kind = head
        self.ops, self.module, self.dtype, self.device = ops, module, dtype, torch.device(device)
        self.pre = enc_prefix
        self.B, self.N, self.H, self.W, self.K = B, B, H, W, num_classes
        self.Cc, self.D, self.depth, self.heads, self.dh, self.mlp = (cfg[k] for k in ("channels", "dim", "depth", "heads", "dim_head", "mlp_dim"))
        assert cfg["patch_size"] == 16 and H % 16 == 0 and W % 16 == 0 and H == W == cfg["image_size"]
        assert self.dh == 64, "the attention kernels are built for dim_head == 64 (vision_transformer.py:103 default)"
        assert self.D % 32 == 0 and self.mlp % 32 == 0
        self.G = H // 16
        self.T = self.G * self.G + 1
        self.Tp = (self.T + 15) // 16 * 16
        self.R = B * self.Tp
        self.inner = self.heads * self.dh
        self.PD = 256 * self.Cc
        self.scale = self.dh ** -0.5
        self.conv_impl = conv_impl
        self.params = FlatParams(module)
        self._alloc()

    # ------------------------------------------------------------------------------------------
    def _mat(self, C, dtype=None):
        return torch.zeros(self.R, C, dtype=dtype or self.dtype, device=self.device)

    def _vec(self):
        return torch.zeros(self.R, dtype=torch.float32, device=self.device)

    def _alloc(self):
        D, dev = self.D, self.device
        self.a0, self.mp, self.rp = self._mat(self.PD), self._vec(), self._vec()        # patchify + LN(patch_dim)
        self.e1, self.e2, self.me, self.re = self._mat(D), self._mat(D), self._vec(), self._vec()
        self.blocks: List[_Block] = []
        x = self._mat(D)
        self.x0 = x
        for l in range(self.depth):
            b = _Block()
            b.xa, b.xn1, b.m1, b.r1 = x, self._mat(D), self._vec(), self._vec()
            b.qkv, b.att = self._mat(3 * self.inner), self._mat(self.inner)
            b.probs = torch.zeros(self.B * self.heads * self.Tp * self.Tp, dtype=self.dtype, device=dev)
            b.xm, b.xn2, b.m2, b.r2 = self._mat(D), self._mat(D), self._vec(), self._vec()
            b.u, b.h = self._mat(self.mlp), self._mat(self.mlp)
            b.pa, b.pf = f"{self.pre}transformer.layers.{l}.0", f"{self.pre}transformer.layers.{l}.1.net"
            x = self._mat(D)
            self.blocks.append(b)
        self.xL = x
        self.tok, self.mf, self.rf = self._mat(D), self._vec(), self._vec()
        self.z = self._mat(HEAD_PAD)
        self.logits = torch.zeros(self.B, self.K, self.H, self.W, dtype=torch.float32, device=dev)
        # gradient scratch (shared by all layers)
        self.dx, self.dxn = self._mat(D), self._mat(D)
        self.dqkv, self.datt = self._mat(3 * self.inner), self._mat(self.inner)
        self.ds = torch.zeros(self.B * self.heads * self.Tp * self.Tp, dtype=self.dtype, device=dev)
        self.dh_ = self._mat(self.mlp)
        self.dz = self._mat(HEAD_PAD)
        self.da0 = self._mat(self.PD)
        self.de1 = self._mat(D)
        # Linear layers: name -> (out, in); packed copies in the storage dtype: [out][in] (forward) and [in][out] (data gradient)
        self.linears: Dict[str, tuple] = {f"{self.pre}to_patch_embedding.2": (D, self.PD)}
        for b in self.blocks:
            self.linears[f"{b.pa}.to_qkv"] = (3 * self.inner, D)
            self.linears[f"{b.pa}.to_out.0"] = (D, self.inner)
            self.linears[f"{b.pf}.1"] = (self.mlp, D)
            self.linears[f"{b.pf}.4"] = (D, self.mlp)
        self.wp: Dict[str, torch.Tensor] = {}
        for name, (o, i) in self.linears.items():
            self.wp[f"{name}.fwd"] = torch.zeros(o * i, dtype=self.dtype, device=dev)
            self.wp[f"{name}.dgrad"] = torch.zeros(o * i, dtype=self.dtype, device=dev)
        self.wp["head.fwd"] = torch.zeros(HEAD_PAD * D, dtype=self.dtype, device=dev)     # rows >= K stay zero
        self.wp["head.dgrad"] = torch.zeros(D * HEAD_PAD, dtype=self.dtype, device=dev)
        self.wp["head.bias"] = torch.zeros(HEAD_PAD, dtype=torch.float32, device=dev)
        self.gp_head = torch.zeros(HEAD_PAD * D, dtype=torch.float32, device=dev)
        self.gp_head_bias = torch.zeros(HEAD_PAD, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------------------------------
    def _pack_jobs(self):
        P, jobs = self.params, []
        for name, (o, i) in self.linears.items():
            w = P.p(f"{name}.weight")
            jobs.append((w, self.wp[f"{name}.fwd"], (o * i,), (1,), 0))
            jobs.append((w, self.wp[f"{name}.dgrad"], (i, o), (1, i), 0))                     # [in][out] = w[out][in]
        D, K = self.D, self.K
        w = P.p("head.weight")                                                              # (K, D, 1, 1)
        jobs.append((w, self.wp["head.fwd"], (K * D,), (1,), 0))
        jobs.append((w, self.wp["head.dgrad"], (D, K), (1, D), 0, (HEAD_PAD, 1), 0))
        jobs.append((P.p("head.bias"), self.wp["head.bias"], (K,), (1,), 0))
        return jobs

    def _unpack_jobs(self):
        P = self.params
        return [(self.gp_head, P.g("head.weight"), (self.K * self.D,), (1,), 0),
                (self.gp_head_bias, P.g("head.bias"), (self.K,), (1,), 0)]

    def _tables(self):
        key = (self.params.flat.data_ptr(), self.params.grad.data_ptr())
        if getattr(self, "_table_key", None) != key:
            self._pack_table = self.ops.make_permute_table(self._pack_jobs(), self.device)
            self._unpack_table = self.ops.make_permute_table(self._unpack_jobs(), self.device)
            self._table_key = key
        return self._pack_table, self._unpack_table

    # ------------------------------------------------------------------------------------------
    def _linear(self, a: torch.Tensor, name: str, out: torch.Tensor, bias: Optional[torch.Tensor], accumulate: bool = False):
        self.ops.conv2d(1, self.R // 16, 16, 1, [tok_view(a)], self.wp[f"{name}.fwd"], bias, [tok_view(out)], [accumulate], None, self.conv_impl)

    def _linear_bwd(self, a: torch.Tensor, name: str, dy: torch.Tensor, da: Optional[torch.Tensor], has_bias: bool):
        """dW (assigned, straight into the flat gradient), db, and (optionally) da = dy W."""
        ops, P = self.ops, self.params
        ops.conv2d_wgrad(1, self.R // 16, 16, 1, [tok_view(a)], [tok_view(dy)], P.g(f"{name}.weight"), False, self.conv_impl)
        if has_bias:
            self._colsum(dy, P.g(f"{name}.bias"))
        if da is not None:
            ops.conv2d(1, self.R // 16, 16, 1, [tok_view(dy)], self.wp[f"{name}.dgrad"], None, [tok_view(da)], [False], None, self.conv_impl)

    def _colsum(self, m: torch.Tensor, out: torch.Tensor):
        v = tok_view(m)
        for c0 in range(0, v.C, 1024):
            c = min(1024, v.C - c0)
            self.ops.channel_sum(v.ch(c0, c), out[c0:c0 + c], True)     # accumulate into the flat gradient, zeroed at the start of backward: no memset launch per bias

    # ------------------------------------------------------------------------------------------
    def forward(self, img: torch.Tensor, training: bool = True) -> torch.Tensor:
        self._encoder_forward(img)
        return self._head_forward(training)

    def _encoder_forward(self, img: torch.Tensor):
        ops, P, pre, B = self.ops, self.params, self.pre, self.B
        assert tuple(img.shape) == (B, self.Cc, self.H, self.W), f"engine was planned for {(B, self.Cc, self.H, self.W)}, got {tuple(img.shape)}"
        P.ensure(self.device)
        img = img.contiguous()
        if img.dtype != torch.float32:
            img = img.float()
        self._img = img
        ops.permute_cast_table(self._tables()[0])
        pe = f"{pre}to_patch_embedding"
        ops.patchify_ln(img, self.Tp, P.p(f"{pe}.1.weight"), P.p(f"{pe}.1.bias"), LN_EPS, self.a0, self.mp, self.rp)
        self._linear(self.a0, f"{pe}.2", self.e1, P.p(f"{pe}.2.bias"))
        ops.layernorm_fwd(self.e1, P.p(f"{pe}.3.weight"), P.p(f"{pe}.3.bias"), LN_EPS, self.e2, self.me, self.re, None)
        ops.vit_assemble(B, self.T, self.Tp, self.e2, P.p(f"{pre}cls_token"), P.p(f"{pre}pos_embedding"), self.x0)
        for i, b in enumerate(self.blocks):
            nxt = self.blocks[i + 1].xa if i + 1 < self.depth else self.xL
            ops.layernorm_fwd(b.xa, P.p(f"{b.pa}.norm.weight"), P.p(f"{b.pa}.norm.bias"), LN_EPS, b.xn1, b.m1, b.r1, b.xm)
            self._linear(b.xn1, f"{b.pa}.to_qkv", b.qkv, None)
            ops.attention_fwd(B, self.T, self.Tp, self.heads, self.dh, b.qkv, self.scale, b.att, b.probs)
            self._linear(b.att, f"{b.pa}.to_out.0", b.xm, P.p(f"{b.pa}.to_out.0.bias"), accumulate=True)        # x = attn(x) + x
            ops.layernorm_fwd(b.xm, P.p(f"{b.pf}.0.weight"), P.p(f"{b.pf}.0.bias"), LN_EPS, b.xn2, b.m2, b.r2, nxt)
            self._linear(b.xn2, f"{b.pf}.1", b.u, P.p(f"{b.pf}.1.bias"))
            ops.gelu_fwd(b.u, b.h)
            self._linear(b.h, f"{b.pf}.4", nxt, P.p(f"{b.pf}.4.bias"), accumulate=True)                          # x = ff(x) + x
        ops.layernorm_fwd(self.xL, P.p(f"{pre}transformer.norm.weight"), P.p(f"{pre}transformer.norm.bias"), LN_EPS, self.tok, self.mf, self.rf, None)

    def _head_forward(self, training: bool) -> torch.Tensor:
        ops, B = self.ops, self.B
        # head: the 1x1 conv commutes with the bilinear interpolation -> classify the G x G grid, upsample K planes
        ops.conv2d(1, self.R // 16, 16, 1, [tok_view(self.tok)], self.wp["head.fwd"], self.wp["head.bias"], [tok_view(self.z)], [False], None, self.conv_impl)
        ops.bilinear_up_fwd(B, self.G, self.Tp, 1, self.K, self.H, self.W, self.z, self.logits)
        return self.logits

    def tokens(self) -> torch.Tensor:
        """Encoder output of the last forward as the reference returns it: x[:, 1:] -> [B, T-1, D] fp32."""
        return self.tok.view(self.B, self.Tp, self.D)[:, 1:self.T].float()

    # ------------------------------------------------------------------------------------------
    def backward(self, dlogits: torch.Tensor):
        self.ops.zero_(self.params.grad)        # LayerNorm / attention parameter gradients are accumulated with atomics
        self._head_backward(dlogits)            # leaves d(tokens) in self.dxn (all rows; 0 for the cls / padding rows)
        self.ops.permute_cast_table(self._tables()[1])      # packed head gradients -> flat gradient: the head's slice is final now
        self._grads_ready(self._head_lo(), self.params.numel)
        if getattr(self, "freeze_encoder", False):
            # linear_eval (models/model_utilities.py:160-161: encoder parameters have requires_grad=False): no encoder data- or
            # weight-gradient is computed at all; the encoder slices of the flat gradient stay zero, so Adam leaves them untouched
            return
        self._encoder_backward()

    def _head_lo(self) -> int:
        """First flat-gradient element that belongs to the head (parameters registered after the encoder's)."""
        offs = self.params.offsets
        head = min((off for n, (off, _) in offs.items() if not n.startswith(self.pre)), default=self.params.numel)
        enc_end = max((off + shape.numel() for n, (off, shape) in offs.items() if n.startswith(self.pre)), default=0)
        return head if enc_end <= head else self.params.numel       # head registered before the encoder: no early head bucket

    def _block_range(self, bi: int):
        lo = self.params.offsets[f"{self.blocks[bi].pa}.norm.weight"][0]
        hi = self.params.offsets[f"{self.blocks[bi + 1].pa}.norm.weight"][0] if bi + 1 < self.depth else self._head_lo()
        return lo, hi

    def _inject(self, bi: int):
        """Hook: add gradients that enter the residual stream AFTER block `bi` (multi-level heads); none for the linear head."""

    def _head_backward(self, dlogits: torch.Tensor):
        ops, P, pre, B = self.ops, self.params, self.pre, self.B
        ops.bilinear_up_bwd(B, self.G, self.Tp, 1, self.K, self.H, self.W, dlogits, self.dz)
        ops.conv2d_wgrad(1, self.R // 16, 16, 1, [tok_view(self.tok)], [tok_view(self.dz)], self.gp_head, False, self.conv_impl)
        ops.channel_sum(tok_view(self.dz), self.gp_head_bias, False)
        ops.conv2d(1, self.R // 16, 16, 1, [tok_view(self.dz)], self.wp["head.dgrad"], None, [tok_view(self.dxn)], [False], None, self.conv_impl)

    def _encoder_backward(self):
        ops, P, pre, B = self.ops, self.params, self.pre, self.B
        ops.layernorm_bwd(self.dxn, self.xL, self.mf, self.rf, P.p(f"{pre}transformer.norm.weight"), self.dx, False,
                          P.g(f"{pre}transformer.norm.weight"), P.g(f"{pre}transformer.norm.bias"))
        for bi in reversed(range(self.depth)):
            b = self.blocks[bi]
            self._inject(bi)
            # x_out = xm + W2 gelu(W1 LN2(xm) + b1) + b2
            self._linear_bwd(b.h, f"{b.pf}.4", self.dx, self.dh_, True)
            ops.gelu_bwd(b.u, self.dh_, self.dh_)
            self._linear_bwd(b.xn2, f"{b.pf}.1", self.dh_, self.dxn, True)
            ops.layernorm_bwd(self.dxn, b.xm, b.m2, b.r2, P.p(f"{b.pf}.0.weight"), self.dx, True, P.g(f"{b.pf}.0.weight"), P.g(f"{b.pf}.0.bias"))
            # xm = xa + Wo attention(Wqkv LN1(xa)) + bo
            self._linear_bwd(b.att, f"{b.pa}.to_out.0", self.dx, self.datt, True)
            ops.attention_bwd(B, self.T, self.Tp, self.heads, self.dh, b.qkv, b.probs, self.datt, self.scale, self.dqkv, self.ds)
            self._linear_bwd(b.xn1, f"{b.pa}.to_qkv", self.dqkv, self.dxn, False)
            ops.layernorm_bwd(self.dxn, b.xa, b.m1, b.r1, P.p(f"{b.pa}.norm.weight"), self.dx, True, P.g(f"{b.pa}.norm.weight"), P.g(f"{b.pa}.norm.bias"))
            self._grads_ready(*self._block_range(bi))       # this block's parameter gradients are final: its bucket may leave
        pe = f"{pre}to_patch_embedding"
        ops.vit_assemble_bwd(B, self.T, self.Tp, self.dx, self.dxn, P.g(f"{pre}cls_token"), P.g(f"{pre}pos_embedding"))
        ops.layernorm_bwd(self.dxn, self.e1, self.me, self.re, P.p(f"{pe}.3.weight"), self.de1, False, P.g(f"{pe}.3.weight"), P.g(f"{pe}.3.bias"))
        self._linear_bwd(self.a0, f"{pe}.2", self.de1, self.da0, True)
        ops.patchify_ln_bwd(self._img, self.Tp, self.mp, self.rp, self.da0, P.g(f"{pe}.1.weight"), P.g(f"{pe}.1.bias"))

    def grid_view(self, t: torch.Tensor) -> View:
        """The patch tokens (cls dropped) of a [R, C] token matrix as the NHWC map [B, G, G, C] - a strided view, no copy."""
        Cn = t.shape[1]
        return View(t.view(-1), Cn, self.B, self.G, self.G, Cn, self.Tp * Cn, self.G * Cn, Cn)
