"""ctypes binding of libkurosiwo_b200.so (the C ABI in include/kurosiwo_b200.h).

There is NO CPU fallback: `load()` raises if the shared library is missing, and every op
raises `KsError` on a non-zero status.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence

import torch

KS_F32, KS_BF16 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
_LIB_PATH = Path(__file__).resolve().parent / "libkurosiwo_b200.so"
_lib = None


class KsError(RuntimeError):
    pass


class ks_view_t(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64),
                ("C", C.c_int32), ("_pad", C.c_int32)]


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return KS_F32
    if dt == torch.bfloat16:
        return KS_BF16
    raise KsError(f"unsupported storage dtype {dt}")


@dataclass
class View:
    """Strided NHWC view into a flat storage tensor (channel stride 1)."""
    base: torch.Tensor
    offset: int
    N: int
    H: int
    W: int
    C: int
    sn: int
    sh: int
    sw: int

    @staticmethod
    def alloc(N, H, W, C, dtype, device, zero=True) -> "View":
        t = (torch.zeros if zero else torch.empty)(N * H * W * C, dtype=dtype, device=device)
        return View(t, 0, N, H, W, C, H * W * C, W * C, C)

    def ch(self, c0: int, c: int) -> "View":
        assert 0 <= c0 and c0 + c <= self.C
        return View(self.base, self.offset + c0, self.N, self.H, self.W, c, self.sn, self.sh, self.sw)

    def phase(self, i: int, j: int) -> "View":
        """2x2-strided phase (i,j): pixels (2h+i, 2w+j)."""
        return View(self.base, self.offset + i * self.sh + j * self.sw, self.N, self.H // 2, self.W // 2, self.C,
                    self.sn, 2 * self.sh, 2 * self.sw)

    def tensor(self) -> torch.Tensor:
        return torch.as_strided(self.base, (self.N, self.H, self.W, self.C), (self.sn, self.sh, self.sw, 1), self.offset)

    def c_view(self) -> ks_view_t:
        return ks_view_t(self.base.data_ptr() + self.offset * self.base.element_size(), self.sn, self.sh, self.sw, self.C, 0)

    @property
    def dtype(self):
        return self.base.dtype


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise KsError(f"{_LIB_PATH} is missing: run `python -m kurosiwo_b200.build` (nvcc, sm_100a). "
                      "There is no CPU fallback.")
    lib = C.CDLL(str(_LIB_PATH))
    lib.ks_error_string.restype = C.c_char_p
    lib.ks_ce_dice_workspace_bytes.restype = C.c_int64
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "ks_version", "ks_error_string", "ks_set_option", "ks_reset_options", "ks_permute_cast", "ks_permute_cast_batched", "ks_conv2d", "ks_conv2d_wgrad", "ks_conv2d_wgrad_bias", "ks_stem_conv3x3", "ks_stem_wgrad3x3",
    "ks_bn_stats", "ks_bn_finalize", "ks_bn_act", "ks_bn_bwd_reduce", "ks_bn_bwd_apply", "ks_maxpool2x2_bwd",
    "ks_channel_sum", "ks_ecam_pool", "ks_ecam_gates", "ks_ecam_final", "ks_ecam_bwd_reduce", "ks_ecam_gates_bwd",
    "ks_ecam_bwd_apply", "ks_ce_dice_workspace_bytes", "ks_ce_dice_fwd_bwd", "ks_ce_dice_fwd_bwd_ex", "ks_adam_step", "ks_adamw_step", "ks_sgd_step",
    "ks_softmax_head_fwd", "ks_softmax_head_bwd", "ks_dropout_mask", "ks_channel_scale", "ks_absdiff_fwd", "ks_absdiff_bwd",
    "ks_layernorm_fwd", "ks_layernorm_bwd", "ks_patchify_ln", "ks_patchify_ln_bwd", "ks_vit_assemble", "ks_vit_assemble_bwd",
    "ks_confusion_update", "ks_confusion_update_grouped", "ks_sar_preprocess", "ks_attention_fwd", "ks_attention_bwd", "ks_conv2d_strided", "ks_conv2d_strided_dgrad", "ks_conv2d_strided_wgrad", "ks_im2col", "ks_col2im",
    "ks_xattention_fwd", "ks_xattention_bwd", "ks_dwconv3x3_fwd", "ks_dwconv3x3_bwd", "ks_bilinear_nhwc_fwd", "ks_bilinear_nhwc_bwd",
    "ks_relu_fwd", "ks_relu_bwd", "ks_sigmoid_head_fwd", "ks_sigmoid_head_bwd", "ks_dropout_apply", "ks_branch_add", "ks_branch_scale", "ks_gelu_fwd", "ks_gelu_bwd", "ks_bilinear_up_fwd", "ks_bilinear_up_bwd", "ks_adaptive_avgpool_fwd", "ks_adaptive_avgpool_bwd",
]


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _views(vs: Sequence[View]):
    arr = (ks_view_t * len(vs))(*[v.c_view() for v in vs])
    return arr


def _vp(v: Optional[View]):
    return None if v is None else C.byref(v.c_view())


class CudaOps:
    """The product op set: every method is one C-ABI call on torch's current CUDA stream."""

    name = "cuda"

    def __init__(self):
        self.lib = load()
        if not torch.cuda.is_available():
            raise KsError("kurosiwo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.launches = 0

    # -- helpers -----------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _check(self, rc: int, what: str):
        self.launches += 1
        if rc != 0:
            raise KsError(f"{what} failed: {self.lib.ks_error_string(C.c_int(rc)).decode()} (code {rc})")

    def reset_options(self):
        """All perf-experiment / A-B switches back to their defaults (they are process-global state of the library)."""
        self._check(self.lib.ks_reset_options(), "ks_reset_options")

    def set_option(self, name: str, value: int):
        rc = self.lib.ks_set_option(name.encode(), C.c_int(value))
        if rc != 0:
            raise KsError(f"unknown option {name}")

    # -- plumbing ----------------------------------------------------------------------------
    def permute_cast(self, src: torch.Tensor, dst: torch.Tensor, dims, strides, accumulate=False, src_offset=0):
        d = list(dims) + [1] * (4 - len(dims))
        s = list(strides) + [0] * (4 - len(strides))
        sp = C.c_void_p(src.data_ptr() + src_offset * src.element_size())
        rc = self.lib.ks_permute_cast(dtype_code(src.dtype), sp, dtype_code(dst.dtype), _p(dst),
                                      *[C.c_int(x) for x in d], *[C.c_int64(x) for x in s],
                                      C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_permute_cast")

    def make_permute_table(self, jobs, device):
        """jobs: list of (src, dst, dims, strides, src_offset[, dst_strides, dst_offset[, scale]]) -> opaque table for
        permute_cast_table (built once).  Without dst_strides the destination is written contiguously."""
        import numpy as np
        rec = np.zeros(len(jobs), dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("total", "<i8"), ("s", "<i8", 4),
                                                  ("d", "<i4", 3), ("sdt", "<i4"), ("ddt", "<i4"), ("acc", "<i4"),
                                                  ("t", "<i8", 4), ("dstr", "<i4"), ("scale", "<f4")], align=True))
        assert rec.dtype.itemsize == 120, rec.dtype.itemsize
        chunks = []
        for i, job in enumerate(jobs):
            src, dst, dims, strides, off = job[:5]
            d = list(dims) + [1] * (4 - len(dims))
            st = list(strides) + [0] * (4 - len(strides))
            total = d[0] * d[1] * d[2] * d[3]
            dstr, doff, strided = [0, 0, 0, 0], 0, 0
            if len(job) > 5 and job[5] is not None:
                dstr = list(job[5]) + [0] * (4 - len(job[5]))
                doff = job[6] if len(job) > 6 else 0
                strided = 1
            scale = float(job[7]) if len(job) > 7 and job[7] is not None else 0.0          # 0 = no scaling
            # 2-D transposes with a unit source stride along the SLOW destination dimension (the [in][out] data-gradient copies of
            # Linear weights): tiled through shared memory (dst_strided = 2, chunk = 64 x 64 tile index)
            tiled = (not strided and d[2] == 1 and d[3] == 1 and st[0] == 1 and st[1] >= d[0] and d[0] >= 64 and d[1] >= 64)
            if tiled:
                strided = 2
            rec[i] = (src.data_ptr() + off * src.element_size(), dst.data_ptr() + doff * dst.element_size(), total, st, d[1:],
                      dtype_code(src.dtype), dtype_code(dst.dtype), 0, dstr, strided, scale)
            if tiled:
                chunks += [(i, t) for t in range(((d[0] + 63) // 64) * ((d[1] + 63) // 64))]
            else:
                chunks += [(i, c0) for c0 in range(0, total, 4096)]
        jobs_dev = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).to(device)
        chunks_dev = torch.tensor(chunks, dtype=torch.int32, device=device).reshape(-1)
        return (jobs_dev, chunks_dev, len(chunks), [j[:2] for j in jobs])

    def permute_cast_table(self, table):
        jobs_dev, chunks_dev, n, _keep = table
        rc = self.lib.ks_permute_cast_batched(_p(jobs_dev), _p(chunks_dev), C.c_int(n), self._stream())
        self._check(rc, "ks_permute_cast_batched")

    # -- convolution -------------------------------------------------------------------------
    def conv2d(self, N, H, W, ksize, srcs, weight, bias, dsts, acc=None, stats=None, impl=IMPL_AUTO):
        acc = acc or [False] * len(dsts)
        accs = (C.c_int * len(dsts))(*[int(a) for a in acc])
        rc = self.lib.ks_conv2d(dtype_code(srcs[0].dtype), C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(ksize),
                                _views(srcs), C.c_int(len(srcs)), _p(weight), _p(bias),
                                _views(dsts), C.c_int(len(dsts)), accs, _p(stats), C.c_int(impl), self._stream())
        self._check(rc, "ks_conv2d")

    def conv2d_wgrad(self, N, H, W, ksize, xs, dys, dw, accumulate=False, impl=IMPL_AUTO):
        rc = self.lib.ks_conv2d_wgrad(dtype_code(xs[0].dtype), C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(ksize),
                                      _views(xs), C.c_int(len(xs)), _views(dys), C.c_int(len(dys)), _p(dw),
                                      C.c_int(int(accumulate)), C.c_int(impl), self._stream())
        self._check(rc, "ks_conv2d_wgrad")

    def conv2d_wgrad_bias(self, N, H, W, ksize, xs, dys, dw, dbias, bias_mod=0, accumulate=False, accumulate_bias=False, impl=IMPL_AUTO):
        rc = self.lib.ks_conv2d_wgrad_bias(dtype_code(xs[0].dtype), C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(ksize),
                                           _views(xs), C.c_int(len(xs)), _views(dys), C.c_int(len(dys)), _p(dw), C.c_int(int(accumulate)),
                                           _p(dbias), C.c_int(bias_mod), C.c_int(int(accumulate_bias)), C.c_int(impl), self._stream())
        self._check(rc, "ks_conv2d_wgrad_bias")

    def stem_conv3x3(self, x_nchw: torch.Tensor, w_oihw, bias, dst: View, stats=None):
        N, Cin, H, W = x_nchw.shape
        rc = self.lib.ks_stem_conv3x3(dtype_code(dst.dtype), C.c_int(N), C.c_int(Cin), C.c_int(H), C.c_int(W), _p(x_nchw), _p(w_oihw),
                                      _p(bias), _vp(dst), _p(stats), self._stream())
        self._check(rc, "ks_stem_conv3x3")

    def stem_wgrad3x3(self, x_nchw: torch.Tensor, dy: View, dw_oihw, accumulate=False):
        N, Cin, H, W = x_nchw.shape
        rc = self.lib.ks_stem_wgrad3x3(dtype_code(dy.dtype), C.c_int(N), C.c_int(Cin), C.c_int(H), C.c_int(W), _p(x_nchw), _vp(dy),
                                       _p(dw_oihw), C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_stem_wgrad3x3")

    # -- batch norm / pooling ----------------------------------------------------------------
    def bn_stats(self, x: View, sums):
        rc = self.lib.ks_bn_stats(dtype_code(x.dtype), C.c_int(x.N), C.c_int(x.H), C.c_int(x.W), _vp(x), _p(sums), self._stream())
        self._check(rc, "ks_bn_stats")

    def bn_finalize(self, Cn, count, sums, gamma, beta, eps, momentum, rmean, rvar, scale, shift, mean, rstd):
        rc = self.lib.ks_bn_finalize(C.c_int(Cn), C.c_double(count), _p(sums), _p(gamma), _p(beta), C.c_float(eps),
                                     C.c_float(momentum), _p(rmean), _p(rvar), _p(scale), _p(shift), _p(mean), _p(rstd),
                                     self._stream())
        self._check(rc, "ks_bn_finalize")

    def bn_act(self, y: View, scale, shift, res: Optional[View], relu: bool, out: View, pool: Optional[View]):
        rc = self.lib.ks_bn_act(dtype_code(y.dtype), C.c_int(y.N), C.c_int(y.H), C.c_int(y.W), _vp(y), _p(scale), _p(shift),
                                _vp(res), C.c_int(int(relu)), _vp(out), _vp(pool), self._stream())
        self._check(rc, "ks_bn_act")

    def bn_bwd_reduce(self, dout: View, out: Optional[View], y: View, scale, shift, mean, rstd, sums, dpool: Optional[View] = None):
        rc = self.lib.ks_bn_bwd_reduce(dtype_code(y.dtype), C.c_int(y.N), C.c_int(y.H), C.c_int(y.W), _vp(dout), _vp(out), _vp(y),
                                       _vp(dpool), _p(scale), _p(shift), _p(mean), _p(rstd), _p(sums), self._stream())
        self._check(rc, "ks_bn_bwd_reduce")

    def bn_bwd_apply(self, g: View, premasked: bool, y: View, scale, shift, mean, rstd, gamma, sums, count, add: Optional[View],
                     dy: View, dgamma, dbeta, dsum_out, accumulate: bool):
        rc = self.lib.ks_bn_bwd_apply(dtype_code(y.dtype), C.c_int(y.N), C.c_int(y.H), C.c_int(y.W), _vp(g), C.c_int(int(premasked)), _vp(y),
                                      _p(scale), _p(shift), _p(mean), _p(rstd), _p(gamma), _p(sums), C.c_double(count), _vp(add),
                                      _vp(dy), _p(dgamma), _p(dbeta), _p(dsum_out), C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_bn_bwd_apply")

    def maxpool2x2_bwd(self, x: View, dpool: View, dx: View, accumulate: bool):
        rc = self.lib.ks_maxpool2x2_bwd(dtype_code(x.dtype), C.c_int(dpool.N), C.c_int(dpool.H), C.c_int(dpool.W), _vp(x), _vp(dpool),
                                        _vp(dx), C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_maxpool2x2_bwd")

    def channel_sum(self, x: View, out, accumulate: bool):
        rc = self.lib.ks_channel_sum(dtype_code(x.dtype), C.c_int(x.N), C.c_int(x.H), C.c_int(x.W), _vp(x), _p(out),
                                     C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_channel_sum")

    # -- ECAM head ---------------------------------------------------------------------------
    def ecam_pool(self, xs, pooled, argmax, scratch):
        x = xs[0]
        rc = self.lib.ks_ecam_pool(dtype_code(x.dtype), C.c_int(x.N), C.c_int(x.H), C.c_int(x.W), _views(xs), C.c_int(len(xs)),
                                   _p(pooled), _p(argmax), _p(scratch), self._stream())
        self._check(rc, "ks_ecam_pool")

    def ecam_gates(self, N, Cb, J, hid, hid1, pooled, w_fc1, w_fc2, w1_fc1, w1_fc2, gates, hidden):
        rc = self.lib.ks_ecam_gates(C.c_int(N), C.c_int(Cb), C.c_int(J), C.c_int(hid), C.c_int(hid1), _p(pooled), _p(w_fc1),
                                    _p(w_fc2), _p(w1_fc1), _p(w1_fc2), _p(gates), _p(hidden), self._stream())
        self._check(rc, "ks_ecam_gates")

    def ecam_final(self, xs, gates, wf, bf, K, logits, pooled=None, argmax=None):
        x = xs[0]
        rc = self.lib.ks_ecam_final(dtype_code(x.dtype), C.c_int(x.N), C.c_int(x.H), C.c_int(x.W), _views(xs), C.c_int(len(xs)),
                                    _p(gates), _p(wf), _p(bf), C.c_int(K), _p(logits), _p(pooled), _p(argmax), self._stream())
        self._check(rc, "ks_ecam_final")

    def ecam_bwd_reduce(self, xs, K, dlogits, red):
        x = xs[0]
        rc = self.lib.ks_ecam_bwd_reduce(dtype_code(x.dtype), C.c_int(x.N), C.c_int(x.H), C.c_int(x.W), _views(xs), C.c_int(len(xs)),
                                         C.c_int(K), _p(dlogits), _p(red), self._stream())
        self._check(rc, "ks_ecam_bwd_reduce")

    def ecam_gates_bwd(self, N, Cb, J, hid, hid1, K, pooled, hidden, gates, red, wf, w_fc1, w_fc2, w1_fc1, w1_fc2,
                       dpooled, dwf, dbf, dw_fc1, dw_fc2, dw1_fc1, dw1_fc2, accumulate=False):
        rc = self.lib.ks_ecam_gates_bwd(C.c_int(N), C.c_int(Cb), C.c_int(J), C.c_int(hid), C.c_int(hid1), C.c_int(K), _p(pooled),
                                        _p(hidden), _p(gates), _p(red), _p(wf), _p(w_fc1), _p(w_fc2), _p(w1_fc1), _p(w1_fc2),
                                        _p(dpooled), _p(dwf), _p(dbf), _p(dw_fc1), _p(dw_fc2), _p(dw1_fc1), _p(dw1_fc2),
                                        C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_ecam_gates_bwd")

    def ecam_bwd_apply(self, dxs, gates, wf, K, dlogits, dpooled, argmax):
        x = dxs[0]
        rc = self.lib.ks_ecam_bwd_apply(dtype_code(x.dtype), C.c_int(x.N), C.c_int(x.H), C.c_int(x.W), C.c_int(len(dxs)), C.c_int(x.C),
                                        _p(gates), _p(wf), C.c_int(K), _p(dlogits), _p(dpooled), _p(argmax), _views(dxs),
                                        self._stream())
        self._check(rc, "ks_ecam_bwd_apply")

    # -- Siamese U-Net passes ------------------------------------------------------------------
    def softmax_head_fwd(self, z: View, K: int, log_mode: bool, out: torch.Tensor):
        rc = self.lib.ks_softmax_head_fwd(dtype_code(z.dtype), C.c_int(z.N), C.c_int(z.H), C.c_int(z.W), _vp(z), C.c_int(K),
                                          C.c_int(int(log_mode)), _p(out), self._stream())
        self._check(rc, "ks_softmax_head_fwd")

    def softmax_head_bwd(self, out: torch.Tensor, dout: torch.Tensor, K: int, log_mode: bool, dz: View):
        rc = self.lib.ks_softmax_head_bwd(dtype_code(dz.dtype), C.c_int(dz.N), C.c_int(dz.H), C.c_int(dz.W), _p(out), _p(dout),
                                          C.c_int(K), C.c_int(int(log_mode)), _vp(dz), self._stream())
        self._check(rc, "ks_softmax_head_bwd")

    def dropout_mask(self, mask: torch.Tensor, p: float, seed: int, step: Optional[torch.Tensor]):
        rc = self.lib.ks_dropout_mask(_p(mask), C.c_int64(mask.numel()), C.c_float(p), C.c_uint64(seed & (2 ** 64 - 1)), _p(step),
                                      self._stream())
        self._check(rc, "ks_dropout_mask")

    def channel_scale(self, x: View, m: torch.Tensor):
        rc = self.lib.ks_channel_scale(dtype_code(x.dtype), C.c_int(x.N), C.c_int(x.H), C.c_int(x.W), _vp(x), _p(m), self._stream())
        self._check(rc, "ks_channel_scale")

    def absdiff_fwd(self, a: View, b: View, out: View):
        rc = self.lib.ks_absdiff_fwd(dtype_code(a.dtype), C.c_int(a.N), C.c_int(a.H), C.c_int(a.W), _vp(a), _vp(b), _vp(out),
                                     self._stream())
        self._check(rc, "ks_absdiff_fwd")

    def absdiff_bwd(self, a: View, b: View, g: View, da: View, acc_a: bool, db: View, acc_b: bool):
        rc = self.lib.ks_absdiff_bwd(dtype_code(a.dtype), C.c_int(a.N), C.c_int(a.H), C.c_int(a.W), _vp(a), _vp(b), _vp(g),
                                     _vp(da), C.c_int(int(acc_a)), _vp(db), C.c_int(int(acc_b)), self._stream())
        self._check(rc, "ks_absdiff_bwd")

    # -- ViT encoder / FloodViT head passes (token matrices are 2-D row-major torch tensors) ------------
    def layernorm_fwd(self, x, gamma, beta, eps, y, mean=None, rstd=None, copy_out=None):
        rows, Cn = x.shape
        rc = self.lib.ks_layernorm_fwd(dtype_code(x.dtype), C.c_int64(rows), C.c_int(Cn), _p(x), C.c_int64(x.stride(0)), _p(gamma), _p(beta),
                                       C.c_float(eps), _p(y), C.c_int64(y.stride(0)), _p(mean), _p(rstd), _p(copy_out),
                                       C.c_int64(0 if copy_out is None else copy_out.stride(0)), self._stream())
        self._check(rc, "ks_layernorm_fwd")

    def layernorm_bwd(self, dy, x, mean, rstd, gamma, dx, accumulate_dx, dgamma, dbeta):
        rows, Cn = x.shape
        rc = self.lib.ks_layernorm_bwd(dtype_code(x.dtype), C.c_int64(rows), C.c_int(Cn), _p(dy), C.c_int64(dy.stride(0)), _p(x),
                                       C.c_int64(x.stride(0)), _p(mean), _p(rstd), _p(gamma), _p(dx),
                                       C.c_int64(0 if dx is None else dx.stride(0)), C.c_int(int(accumulate_dx)), _p(dgamma), _p(dbeta),
                                       self._stream())
        self._check(rc, "ks_layernorm_bwd")

    def patchify_ln(self, img, Tp, gamma, beta, eps, out, mean, rstd):
        B, Cc, Hi, Wi = img.shape
        rc = self.lib.ks_patchify_ln(dtype_code(out.dtype), C.c_int(B), C.c_int(Cc), C.c_int(Hi), C.c_int(Wi), C.c_int(Tp), _p(img), _p(gamma),
                                     _p(beta), C.c_float(eps), _p(out), _p(mean), _p(rstd), self._stream())
        self._check(rc, "ks_patchify_ln")

    def patchify_ln_bwd(self, img, Tp, mean, rstd, dy, dgamma, dbeta):
        B, Cc, Hi, Wi = img.shape
        rc = self.lib.ks_patchify_ln_bwd(dtype_code(dy.dtype), C.c_int(B), C.c_int(Cc), C.c_int(Hi), C.c_int(Wi), C.c_int(Tp), _p(img), _p(mean),
                                         _p(rstd), _p(dy), _p(dgamma), _p(dbeta), self._stream())
        self._check(rc, "ks_patchify_ln_bwd")

    def vit_assemble(self, B, T, Tp, e, cls, pos, x0):
        rc = self.lib.ks_vit_assemble(dtype_code(e.dtype), C.c_int(B), C.c_int(T), C.c_int(Tp), C.c_int(e.shape[1]), _p(e), _p(cls), _p(pos),
                                      _p(x0), self._stream())
        self._check(rc, "ks_vit_assemble")

    def vit_assemble_bwd(self, B, T, Tp, dx0, de, dcls, dpos):
        rc = self.lib.ks_vit_assemble_bwd(dtype_code(dx0.dtype), C.c_int(B), C.c_int(T), C.c_int(Tp), C.c_int(dx0.shape[1]), _p(dx0), _p(de),
                                          _p(dcls), _p(dpos), self._stream())
        self._check(rc, "ks_vit_assemble_bwd")

    def attention_fwd(self, B, T, Tp, heads, dh, qkv, scale, out, probs):
        rc = self.lib.ks_attention_fwd(dtype_code(qkv.dtype), C.c_int(B), C.c_int(T), C.c_int(Tp), C.c_int(heads), C.c_int(dh), _p(qkv),
                                       C.c_float(scale), _p(out), _p(probs), self._stream())
        self._check(rc, "ks_attention_fwd")

    def attention_bwd(self, B, T, Tp, heads, dh, qkv, probs, dout, scale, dqkv, ds_scratch):
        rc = self.lib.ks_attention_bwd(dtype_code(qkv.dtype), C.c_int(B), C.c_int(T), C.c_int(Tp), C.c_int(heads), C.c_int(dh), _p(qkv),
                                       _p(probs), _p(dout), C.c_float(scale), _p(dqkv), _p(ds_scratch), self._stream())
        self._check(rc, "ks_attention_bwd")

    def gelu_fwd(self, u, h):
        rc = self.lib.ks_gelu_fwd(dtype_code(u.dtype), C.c_int64(u.numel()), _p(u), _p(h), self._stream())
        self._check(rc, "ks_gelu_fwd")

    def gelu_bwd(self, u, dh, du):
        rc = self.lib.ks_gelu_bwd(dtype_code(u.dtype), C.c_int64(u.numel()), _p(u), _p(dh), _p(du), self._stream())
        self._check(rc, "ks_gelu_bwd")

    def bilinear_up_fwd(self, B, G, Tp, row0, K, Ho, Wo, src, dst):
        rc = self.lib.ks_bilinear_up_fwd(dtype_code(src.dtype), C.c_int(B), C.c_int(G), C.c_int(Tp), C.c_int(row0), C.c_int(src.shape[1]),
                                         C.c_int(K), C.c_int(Ho), C.c_int(Wo), _p(src), _p(dst), self._stream())
        self._check(rc, "ks_bilinear_up_fwd")

    def bilinear_up_bwd(self, B, G, Tp, row0, K, Ho, Wo, ddst, dsrc):
        rc = self.lib.ks_bilinear_up_bwd(dtype_code(dsrc.dtype), C.c_int(B), C.c_int(G), C.c_int(Tp), C.c_int(row0), C.c_int(dsrc.shape[1]),
                                         C.c_int(K), C.c_int(Ho), C.c_int(Wo), _p(ddst), _p(dsrc), self._stream())
        self._check(rc, "ks_bilinear_up_bwd")

    def adaptive_avgpool_fwd(self, src: View, S: int, dst: torch.Tensor):
        rc = self.lib.ks_adaptive_avgpool_fwd(dtype_code(src.dtype), C.c_int(src.N), C.c_int(src.H), C.c_int(src.W), C.c_int(S), _vp(src), _p(dst),
                                              self._stream())
        self._check(rc, "ks_adaptive_avgpool_fwd")

    def adaptive_avgpool_bwd(self, ddst: torch.Tensor, S: int, dsrc: View, accumulate=False):
        rc = self.lib.ks_adaptive_avgpool_bwd(dtype_code(dsrc.dtype), C.c_int(dsrc.N), C.c_int(dsrc.H), C.c_int(dsrc.W), C.c_int(S), _p(ddst),
                                              _vp(dsrc), C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_adaptive_avgpool_bwd")

    # -- ChangeFormer passes ---------------------------------------------------------------------
    def conv2d_strided(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, src: View, weight, bias, dst: View):
        rc = self.lib.ks_conv2d_strided(dtype_code(src.dtype), *[C.c_int(v) for v in (N, Hi, Wi, Ho, Wo, ksize, stride, pad)], _vp(src), _p(weight),
                                        _p(bias), _vp(dst), self._stream())
        self._check(rc, "ks_conv2d_strided")

    def conv2d_strided_dgrad(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, dy: View, weight, dx: View, accumulate=False):
        rc = self.lib.ks_conv2d_strided_dgrad(dtype_code(dy.dtype), *[C.c_int(v) for v in (N, Hi, Wi, Ho, Wo, ksize, stride, pad)], _vp(dy),
                                              _p(weight), _vp(dx), C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_conv2d_strided_dgrad")

    def conv2d_strided_wgrad(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, x: View, dy: View, dw, accumulate=False):
        rc = self.lib.ks_conv2d_strided_wgrad(dtype_code(x.dtype), *[C.c_int(v) for v in (N, Hi, Wi, Ho, Wo, ksize, stride, pad)], _vp(x), _vp(dy),
                                              _p(dw), C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_conv2d_strided_wgrad")

    def im2col(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, x: View, col, Kp):
        rc = self.lib.ks_im2col(dtype_code(x.dtype), *[C.c_int(v) for v in (N, Hi, Wi, Ho, Wo, ksize, stride, pad)], _vp(x), _p(col), C.c_int(Kp),
                                self._stream())
        self._check(rc, "ks_im2col")

    def col2im(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, dcol, Kp, dx: View, accumulate=False):
        rc = self.lib.ks_col2im(dtype_code(dx.dtype), *[C.c_int(v) for v in (N, Hi, Wi, Ho, Wo, ksize, stride, pad)], _p(dcol), C.c_int(Kp), _vp(dx),
                                C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_col2im")

    def xattention_fwd(self, B, Nq, Nk, heads, dh, q, kv, scale, out, probs, pdrop=0.0, seed=0, step=None, site=0):
        rc = self.lib.ks_xattention_fwd(dtype_code(q.dtype), *[C.c_int(v) for v in (B, Nq, Nk, heads, dh)], _p(q), C.c_int64(q.stride(0)), _p(kv),
                                        C.c_int64(kv.stride(0)), C.c_float(scale), _p(out), C.c_int64(out.stride(0)), _p(probs),
                                        C.c_float(pdrop), C.c_uint64(seed), _p(step), C.c_int(site), self._stream())
        self._check(rc, "ks_xattention_fwd")

    def xattention_bwd(self, B, Nq, Nk, heads, dh, q, kv, probs, dout, scale, dq, dkv_f32, pdrop=0.0, seed=0, step=None, site=0):
        rc = self.lib.ks_xattention_bwd(dtype_code(q.dtype), *[C.c_int(v) for v in (B, Nq, Nk, heads, dh)], _p(q), C.c_int64(q.stride(0)), _p(kv),
                                        C.c_int64(kv.stride(0)), _p(probs), _p(dout), C.c_int64(dout.stride(0)), C.c_float(scale), _p(dq),
                                        C.c_int64(dq.stride(0)), _p(dkv_f32), C.c_float(pdrop), C.c_uint64(seed), _p(step), C.c_int(site),
                                        self._stream())
        self._check(rc, "ks_xattention_bwd")

    def dropout_apply(self, x, y, p, seed, step, site):
        rc = self.lib.ks_dropout_apply(dtype_code(x.dtype), C.c_int64(x.numel()), _p(x), _p(y), C.c_float(p), C.c_uint64(seed), _p(step),
                                       C.c_int(site), self._stream())
        self._check(rc, "ks_dropout_apply")

    def branch_add(self, x, t, per_sample, p, droppath, seed, step, site):
        rc = self.lib.ks_branch_add(dtype_code(x.dtype), C.c_int64(x.numel()), C.c_int64(per_sample), _p(x), _p(t), C.c_float(p), _p(droppath),
                                    C.c_uint64(seed), _p(step), C.c_int(site), self._stream())
        self._check(rc, "ks_branch_add")

    def branch_scale(self, dx, dt, per_sample, p, droppath, seed, step, site):
        rc = self.lib.ks_branch_scale(dtype_code(dx.dtype), C.c_int64(dx.numel()), C.c_int64(per_sample), _p(dx), _p(dt), C.c_float(p),
                                      _p(droppath), C.c_uint64(seed), _p(step), C.c_int(site), self._stream())
        self._check(rc, "ks_branch_scale")

    def dwconv3x3_fwd(self, N, H, W, x, w9, bias, y):
        rc = self.lib.ks_dwconv3x3_fwd(dtype_code(x.dtype), C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(x.shape[-1]), _p(x), _p(w9), _p(bias), _p(y),
                                       self._stream())
        self._check(rc, "ks_dwconv3x3_fwd")

    def dwconv3x3_bwd(self, N, H, W, x, dy, w9, dx, dw9, dbias):
        rc = self.lib.ks_dwconv3x3_bwd(dtype_code(x.dtype), C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(x.shape[-1]), _p(x), _p(dy), _p(w9), _p(dx),
                                       _p(dw9), _p(dbias), self._stream())
        self._check(rc, "ks_dwconv3x3_bwd")

    def bilinear_nhwc_fwd(self, N, Hi, Wi, Ho, Wo, src, dst, accumulate=False):
        rc = self.lib.ks_bilinear_nhwc_fwd(dtype_code(src.dtype), *[C.c_int(v) for v in (N, Hi, Wi, Ho, Wo, src.shape[-1])], _p(src), _p(dst),
                                           C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_bilinear_nhwc_fwd")

    def bilinear_nhwc_bwd(self, N, Hi, Wi, Ho, Wo, ddst, dsrc, accumulate=False):
        rc = self.lib.ks_bilinear_nhwc_bwd(dtype_code(ddst.dtype), *[C.c_int(v) for v in (N, Hi, Wi, Ho, Wo, ddst.shape[-1])], _p(ddst), _p(dsrc),
                                           C.c_int(int(accumulate)), self._stream())
        self._check(rc, "ks_bilinear_nhwc_bwd")

    def relu_fwd(self, x, y):
        rc = self.lib.ks_relu_fwd(dtype_code(x.dtype), C.c_int64(x.numel()), _p(x), _p(y), self._stream())
        self._check(rc, "ks_relu_fwd")

    def relu_bwd(self, r, g, dx):
        rc = self.lib.ks_relu_bwd(dtype_code(r.dtype), C.c_int64(r.numel()), _p(r), _p(g), _p(dx), self._stream())
        self._check(rc, "ks_relu_bwd")

    def sigmoid_head_fwd(self, z: View, K: int, out: torch.Tensor):
        rc = self.lib.ks_sigmoid_head_fwd(dtype_code(z.dtype), C.c_int(z.N), C.c_int(z.H), C.c_int(z.W), _vp(z), C.c_int(K), _p(out), self._stream())
        self._check(rc, "ks_sigmoid_head_fwd")

    def sigmoid_head_bwd(self, out: torch.Tensor, dout: torch.Tensor, K: int, dz: View):
        rc = self.lib.ks_sigmoid_head_bwd(dtype_code(dz.dtype), C.c_int(dz.N), C.c_int(dz.H), C.c_int(dz.W), _p(out), _p(dout), C.c_int(K), _vp(dz),
                                          self._stream())
        self._check(rc, "ks_sigmoid_head_bwd")

    # -- loss --------------------------------------------------------------------------------
    def ce_dice_workspace(self, N: int, device) -> torch.Tensor:
        nbytes = int(self.lib.ks_ce_dice_workspace_bytes(C.c_int(N)))
        return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=device)   # zero-filled ONCE; every call leaves it clean

    def ce_dice(self, logits, labels, class_weights, ignore_index, grad_scale, loss_out, dlogits, pred, workspace, dice_weight=1.0):
        N, K = logits.shape[0], logits.shape[1]
        HW = logits.numel() // (N * K)
        rc = self.lib.ks_ce_dice_fwd_bwd_ex(_p(logits), _p(labels), C.c_int(N), C.c_int(K), C.c_int64(HW), _p(class_weights),
                                            C.c_int(ignore_index), C.c_float(grad_scale), C.c_float(dice_weight), _p(loss_out),
                                            _p(dlogits), _p(pred), _p(workspace), self._stream())
        self._check(rc, "ks_ce_dice_fwd_bwd_ex")

    def confusion_update(self, pred: torch.Tensor, labels: torch.Tensor, K: int, ignore_index: int, mat: torch.Tensor):
        rc = self.lib.ks_confusion_update(_p(pred), _p(labels), C.c_int64(labels.numel()), C.c_int(K), C.c_int(ignore_index), _p(mat),
                                          self._stream())
        self._check(rc, "ks_confusion_update")

    def confusion_update_grouped(self, pred, labels, K, ignore_index, mat, key_a=None, mat_a=None, key_b=None, mat_b=None):
        """pred uint8 [B,H,W], labels int64 [B,H,W]; key_*: int32 [B] device tensors; mat_*: int64 [n_keys, K, K]."""
        B = labels.shape[0]
        rc = self.lib.ks_confusion_update_grouped(_p(pred), _p(labels), C.c_int(B), C.c_int64(labels.numel() // B), C.c_int(K),
                                                  C.c_int(ignore_index), _p(key_a), C.c_int(0 if mat_a is None else mat_a.shape[0]),
                                                  _p(key_b), C.c_int(0 if mat_b is None else mat_b.shape[0]), _p(mat), _p(mat_a), _p(mat_b),
                                                  self._stream())
        self._check(rc, "ks_confusion_update_grouped")

    def sar_preprocess(self, raw: torch.Tensor, out: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, clamp_max: float):
        """raw / out: fp32 [B, C, H, W] contiguous (out may be raw); mean / std: fp32 [C] device tensors."""
        B, Cn = raw.shape[0], raw.shape[1]
        rc = self.lib.ks_sar_preprocess(_p(raw), _p(out), C.c_int(B), C.c_int(Cn), C.c_int64(raw.shape[2] * raw.shape[3]), _p(mean), _p(std),
                                        C.c_float(clamp_max if clamp_max is not None else 0.0), self._stream())
        self._check(rc, "ks_sar_preprocess")

    # -- optimizer ---------------------------------------------------------------------------
    def adam_step(self, p, g, m, v, lr, b1, b2, eps, wd, grad_scale, step):
        rc = self.lib.ks_adam_step(_p(p), _p(g), _p(m), _p(v), C.c_int64(p.numel()), C.c_float(lr), C.c_float(b1), C.c_float(b2),
                                   C.c_float(eps), C.c_float(wd), C.c_float(grad_scale), _p(step), self._stream())
        self._check(rc, "ks_adam_step")

    def adamw_step(self, p, g, m, v, lr, b1, b2, eps, wd, grad_scale, step):
        rc = self.lib.ks_adamw_step(_p(p), _p(g), _p(m), _p(v), C.c_int64(p.numel()), C.c_float(lr), C.c_float(b1), C.c_float(b2),
                                    C.c_float(eps), C.c_float(wd), C.c_float(grad_scale), _p(step), self._stream())
        self._check(rc, "ks_adamw_step")

    def sgd_step(self, p, g, buf, lr, momentum, wd, grad_scale):
        rc = self.lib.ks_sgd_step(_p(p), _p(g), _p(buf), C.c_int64(p.numel()), C.c_float(lr), C.c_float(momentum), C.c_float(wd),
                                  C.c_float(grad_scale), self._stream())
        self._check(rc, "ks_sgd_step")

    # -- misc memory ops (plumbing through torch) ----------------------------------------------
    def zero_(self, t: torch.Tensor):
        t.zero_()


_default_ops: Optional[CudaOps] = None


def default_ops() -> CudaOps:
    global _default_ops
    if _default_ops is None:
        _default_ops = CudaOps()
    return _default_ops
