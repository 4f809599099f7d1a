"""Host-side schedule of the SNUNet-ECAM training step over the C-ABI ops.

Reference path being replaced: models/snunet.py:118-153 (SNUNet_ECAM.forward) and its autograd
backward, training/change_detection_trainer.py:136-177 (forward, loss, backward, optimizer step).

Layout in HBM (N = per-GPU batch, f_l = base*2^l, H_l = H >> l), all NHWC in the storage dtype
(bf16 in perf mode, fp32 in parity mode):
  X[l]    one dense "concat" buffer per pyramid level, channel slots [x_l0A | x_l0B | x_l1 | x_l2 ...]:
          every block writes its output straight into its slot, so each `torch.cat` of
          snunet.py:132-144 is the channel PREFIX of X[l] plus the upsampled tensor - never a copy.
  UP[l,j] ConvTranspose2d(k2,s2) output feeding decoder block (l,j); written by four strided
          1x1-conv phases.
  P[l]    2x2 max-pooled encoder outputs, produced by the same pass that applies BN+residual+ReLU.
  Y1/Hh/Y2 per block execution: pre-BN conv1 output (also the residual), post-BN/ReLU hidden,
          pre-BN conv2 output - kept for backward.
  dX/dUP/dP mirror the activations for gradients; producers accumulate (+=) or assign according
          to a static first-writer analysis, so no gradient buffer is ever zero-filled.
Parameters live in ONE flat fp32 buffer (module parameters are views into it) with a matching flat
gradient buffer: the optimizer step and the data-parallel all-reduce are one launch / one message.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .engine_common import FlatParams, TrainStepMixin, _align  # noqa: F401
from .lib import View

DEC_ORDER = [(0, 1), (1, 1), (0, 2), (2, 1), (1, 2), (0, 3), (3, 1), (2, 2), (1, 3), (0, 4)]  # snunet.py:132-144
NSLOTS = [6, 5, 4, 3, 1]
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class _Exec:
    """One execution of a conv_block_nested (the shared encoder blocks execute twice)."""
    __slots__ = ("name", "level", "srcs", "gsrcs", "out", "dout", "pool", "y1", "h", "y2", "bn", "stats", "bstats", "key")


class SNUNetEngine(TrainStepMixin):
    def __init__(self, ops, module: torch.nn.Module, in_ch: int, num_classes: int, base: int,
                 N: int, H: int, W: int, dtype: torch.dtype, device, conv_impl: int = 0, planar_slots: Optional[bool] = None):
        assert H % 16 == 0 and W % 16 == 0, "SNUNet needs H, W divisible by 16 (four 2x2 poolings)"
        assert num_classes == 3, "the fused head/loss kernels are built for num_classes == 3 (configs/config.json:13)"
        self.ops, self.module, self.dtype, self.device = ops, module, dtype, torch.device(device)
        self.in_ch, self.K, self.base, self.N, self.H, self.W = in_ch, num_classes, base, N, H, W
        self.f = [base * (1 << l) for l in range(5)]
        self.hid, self.hid1 = (4 * base) // 16, base // 4
        self.conv_impl = conv_impl
        # slot layout of the per-level activation store: one dense concat buffer (channel slices) or one dense tensor per slot
        import os
        self.planar = bool(int(os.environ.get("KS_PLANAR_SLOTS", "1"))) if planar_slots is None else planar_slots
        # dedicated HBM-bound stem kernels (NCHW fp32 input read directly) when conv0_0 is Cin<=4 -> 32
        self.use_stem = (base == 32 and in_ch <= 4)
        self._x_in = [None, None]
        self.params = FlatParams(module)
        self._alloc()
        self._build_schedule()
        self.packed_version = -1

    # ------------------------------------------------------------------------------------------
    def _hw(self, l):
        return self.H >> l, self.W >> l

    def _buf(self, l, C):
        h, w = self._hw(l)
        return View.alloc(self.N, h, w, C, self.dtype, self.device)

    def slot(self, buf, l: int, k: int) -> View:
        f = self.f[l]
        if self.planar:
            return buf[l][k if l < 4 else 0]
        return buf[l].ch((k if l < 4 else 0) * f, f)

    def prefix(self, buf, l: int, nslots: int) -> List[View]:
        """Views covering slots 0..nslots-1 of level l (the torch.cat operand list minus the upsampled tensor)."""
        if self.planar:
            return [buf[l][k] for k in range(nslots)]
        return [buf[l].ch(0, self.f[l] * nslots)]

    def _alloc(self):
        N, dev, f = self.N, self.device, self.f
        if self.planar:
            self.X = {l: [self._buf(l, f[l]) for _ in range(NSLOTS[l])] for l in range(5)}
            self.dX = {l: [self._buf(l, f[l]) for _ in range(NSLOTS[l])] for l in range(5)}
        else:
            self.X = {l: self._buf(l, f[l] * NSLOTS[l]) for l in range(5)}
            self.dX = {l: self._buf(l, f[l] * NSLOTS[l]) for l in range(5)}
        self.P, self.dP = {}, {}
        for l in range(1, 5):
            for br in (0, 1):
                if l == 4 and br == 0:
                    continue  # conv4_0 is only run on the second image (snunet.py:124)
                self.P[(l, br)] = self._buf(l, f[l - 1])
                self.dP[(l, br)] = self._buf(l, f[l - 1])
        self.UP = {(l, j): self._buf(l, f[l + 1]) for (l, j) in DEC_ORDER}
        self.dUP = {(l, j): self._buf(l, f[l + 1]) for (l, j) in DEC_ORDER}
        self.xin = [View.alloc(N, self.H, self.W, self.in_ch, self.dtype, dev) for _ in range(2)]
        self.dY = {l: self._buf(l, f[l]) for l in range(5)}
        self.dH = {l: self._buf(l, f[l]) for l in range(5)}
        CT = 5 * f[0]
        self.pooled = torch.zeros(N * 2 * CT, dtype=torch.float32, device=dev)
        self.dpooled = torch.zeros(N * 2 * CT, dtype=torch.float32, device=dev)
        self.argmax = torch.zeros(N * CT, dtype=torch.int32, device=dev)
        self.pool_scratch = torch.zeros(N * CT, dtype=torch.int64, device=dev)
        self.gates = torch.zeros(N * CT, dtype=torch.float32, device=dev)
        self.hidden = torch.zeros(N * 2 * (self.hid + self.hid1), dtype=torch.float32, device=dev)
        self.red = torch.zeros(N * (self.K * 4 * f[0] + self.K), dtype=torch.float64, device=dev)
        self.logits = torch.zeros(N, self.K, self.H, self.W, dtype=torch.float32, device=dev)

    def _build_schedule(self):
        f = self.f
        self.execs: List[_Exec] = []
        self.exec_of: Dict[Tuple, _Exec] = {}

        def add(name, level, srcs, gsrcs, out_key, pool_key, key):
            e = _Exec()
            e.name, e.level, e.srcs, e.gsrcs, e.key = name, level, srcs, gsrcs, key
            e.out = self.slot(self.X, level, out_key)
            e.dout = self.slot(self.dX, level, out_key)
            e.pool = self.P.get(pool_key) if pool_key else None
            e.y1, e.h, e.y2 = self._buf(level, f[level]), self._buf(level, f[level]), self._buf(level, f[level])
            e.bn = torch.zeros(8 * f[level], dtype=torch.float32, device=self.device)   # scale1,shift1,mean1,rstd1,scale2,...
            e.stats = None
            self.execs.append(e)
            self.exec_of[key] = e
            return e

        for br in (0, 1):  # encoder A then B, as the reference (running stats update order)
            for l in range(5):
                if l == 4 and br == 0:
                    continue
                src = self.xin[br] if l == 0 else self.P[(l, br)]
                gsrc = None if l == 0 else self.dP[(l, br)]
                pool_key = (l + 1, br) if (l + 1, br) in self.P else None
                add(f"conv{l}_0", l, [src], [gsrc], br, pool_key, ("enc", l, br))
        for (l, j) in DEC_ORDER:
            add(f"conv{l}_{j}", l, self.prefix(self.X, l, j + 1) + [self.UP[(l, j)]], None, 1 + j, None, ("dec", l, j))
        n = len(self.execs)
        fmax = max(self.f)
        self.stats_all = torch.zeros(n * 2 * 2 * fmax, dtype=torch.float64, device=self.device)
        self.bstats_all = torch.zeros(n * 2 * 2 * fmax, dtype=torch.float64, device=self.device)
        for i, e in enumerate(self.execs):
            fl = f[e.level]
            o = i * 4 * fmax
            e.stats = (self.stats_all[o:o + 2 * fl], self.stats_all[o + 2 * fmax:o + 2 * fmax + 2 * fl])
            e.bstats = (self.bstats_all[o:o + 2 * fl], self.bstats_all[o + 2 * fmax:o + 2 * fmax + 2 * fl])

        # packed weights (storage dtype) and packed weight gradients (fp32)
        self.wp: Dict[str, torch.Tensor] = {}
        self.gp: Dict[str, torch.Tensor] = {}
        self.block_names = sorted({e.name for e in self.execs})
        self.block_cin = {}
        for e in self.execs:
            self.block_cin[e.name] = sum(s.C for s in e.srcs)
        for bn_ in self.block_names:
            l = int(bn_[4])
            cin, fl = self.block_cin[bn_], f[l]
            for tag, (co, ci) in (("conv1", (fl, cin)), ("conv2", (fl, fl))):
                self.wp[f"{bn_}.{tag}.fwd"] = torch.zeros(9 * co * ci, dtype=self.dtype, device=self.device)
                self.wp[f"{bn_}.{tag}.dgrad"] = torch.zeros(9 * co * ci, dtype=self.dtype, device=self.device)
                self.gp[f"{bn_}.{tag}"] = torch.zeros(9 * co * ci, dtype=torch.float32, device=self.device)
        self.up_names = {}
        for (l, j) in DEC_ORDER:
            nm = f"Up{l + 1}_{j - 1}"
            self.up_names[(l, j)] = nm
            c = f[l + 1]
            self.wp[f"{nm}.fwd"] = torch.zeros(4 * c * c, dtype=self.dtype, device=self.device)
            self.wp[f"{nm}.dgrad"] = torch.zeros(4 * c * c, dtype=self.dtype, device=self.device)
            self.wp[f"{nm}.bias4"] = torch.zeros(4 * c, dtype=torch.float32, device=self.device)
            self.gp[nm] = torch.zeros(4 * c * c, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------------------------------
    def _pack_jobs(self):
        P, jobs = self.params, []
        for bn_ in self.block_names:
            l = int(bn_[4])
            cin, fl = self.block_cin[bn_], self.f[l]
            for tag, (co, ci) in (("conv1", (fl, cin)), ("conv2", (fl, fl))):
                if self.use_stem and bn_ == "conv0_0" and tag == "conv1":
                    continue
                w = P.p(f"{bn_}.{tag}.weight")  # OIHW
                # fwd  [t][o][i] = w[o][i][t]
                jobs.append((w, self.wp[f"{bn_}.{tag}.fwd"], (9, co, ci), (1, ci * 9, 9), 0))
                # dgrad [t][i][o] = w[o][i][8-t]   (180-degree rotated taps, in/out swapped)
                jobs.append((w, self.wp[f"{bn_}.{tag}.dgrad"], (9, ci, co), (-1, 9, ci * 9), 8))
        for (l, j), nm in self.up_names.items():
            c = self.f[l + 1]
            w = P.p(f"{nm}.up.weight")  # (Cin, Cout, 2, 2)
            jobs.append((w, self.wp[f"{nm}.fwd"], (4, c, c), (1, 4, c * 4), 0))       # [k][co][ci]
            jobs.append((w, self.wp[f"{nm}.dgrad"], (c, 4, c), (c * 4, 1, 4), 0))     # [ci][k][co]
            jobs.append((P.p(f"{nm}.up.bias"), self.wp[f"{nm}.bias4"], (4, c), (0, 1), 0))
        return jobs

    def _unpack_jobs(self):
        """Three groups by the point of the backward at which the packed gradients are final (data-parallel overlap, see backward()):
        "dec" = nested-decoder blocks + every ConvTranspose (after the decoder loop), "l4" = conv4_0 (executed for one date only: final
        after its own backward), "enc" = the shared encoder blocks conv0_0..conv3_0 (final after the second date, i.e. at the end)."""
        P, jobs = self.params, {"dec": [], "l4": [], "enc": []}
        for bn_ in self.block_names:
            l = int(bn_[4])
            cin, fl = self.block_cin[bn_], self.f[l]
            grp = "dec" if int(bn_[6:]) >= 1 else ("l4" if l == 4 else "enc")
            for tag, (co, ci) in (("conv1", (fl, cin)), ("conv2", (fl, fl))):
                if self.use_stem and bn_ == "conv0_0" and tag == "conv1":
                    continue
                # grad[o][i][t] = gp[t][o][i]
                jobs[grp].append((self.gp[f"{bn_}.{tag}"], P.g(f"{bn_}.{tag}.weight"), (co, ci, 9), (ci, 1, co * ci), 0))
        for (l, j), nm in self.up_names.items():
            c = self.f[l + 1]
            # grad[ci][co][k] = gp[k][co][ci]
            jobs["dec"].append((self.gp[nm], P.g(f"{nm}.up.weight"), (c, c, 4), (1, c, c * c), 0))
        return jobs

    def _tables(self):
        """Permute tables hold raw pointers into the flat parameter/gradient buffers: rebuild when those move."""
        key = (self.params.flat.data_ptr(), self.params.grad.data_ptr())
        if getattr(self, "_table_key", None) != key:
            self._pack_table = self.ops.make_permute_table(self._pack_jobs(), self.device)
            self._unpack_table = {g: self.ops.make_permute_table(j, self.device) for g, j in self._unpack_jobs().items() if j}
            self._table_key = key
        return self._pack_table, self._unpack_table

    def _pack_weights(self):
        self.ops.permute_cast_table(self._tables()[0])

    def _unpack_grads(self, group: str):
        t = self._tables()[1].get(group)
        if t is not None:
            self.ops.permute_cast_table(t)

    def _ensure_nbt(self):
        """num_batches_tracked of every BatchNorm as views into one int64 buffer (re-pointed after .to()/load)."""
        mods = [(f"{bn_}.{t}", self.module.get_submodule(f"{bn_}.{t}")) for bn_ in self.block_names for t in ("bn1", "bn2")]
        flat = getattr(self, "nbt_all", None)
        ok = flat is not None and flat.device == self.device and all(
            m.num_batches_tracked.data_ptr() == flat.data_ptr() + 8 * i for i, (_, m) in enumerate(mods))
        if ok:
            return
        flat = torch.zeros(len(mods), dtype=torch.int64, device=self.device)
        incr = torch.zeros(len(mods), dtype=torch.int64, device=self.device)
        execs_per_block = {}
        for e in self.execs:
            execs_per_block[e.name] = execs_per_block.get(e.name, 0) + 1
        for i, (name, m) in enumerate(mods):
            flat[i] = m.num_batches_tracked.to(self.device)
            m._buffers["num_batches_tracked"] = flat[i]
            incr[i] = execs_per_block[name.split(".")[0]]
        self.nbt_all, self.nbt_incr = flat, incr

    def _buf_(self, name: str) -> torch.Tensor:
        mod, _, leaf = name.rpartition(".")
        return getattr(self.module.get_submodule(mod), leaf)

    # ------------------------------------------------------------------------------------------
    def _block_forward(self, e: _Exec, training: bool):
        ops, P, N = self.ops, self.params, self.N
        h_, w_ = self._hw(e.level)
        fl = self.f[e.level]
        nm = e.name
        count = float(N * h_ * w_)
        bn = e.bn
        sc1, sh1, mu1, rs1, sc2, sh2, mu2, rs2 = [bn[i * fl:(i + 1) * fl] for i in range(8)]
        for idx, (tag, bnt, srcs, y, sc, sh, mu, rs) in enumerate((("conv1", "bn1", e.srcs, e.y1, sc1, sh1, mu1, rs1),
                                                                   ("conv2", "bn2", [e.h], e.y2, sc2, sh2, mu2, rs2))):
            stats = e.stats[idx] if training else None
            if idx == 0 and self.use_stem and e.key[0] == "enc" and e.level == 0:
                ops.stem_conv3x3(self._x_in[e.key[2]], P.p(f"{nm}.conv1.weight"), P.p(f"{nm}.conv1.bias"), y, stats)
            else:
                ops.conv2d(N, h_, w_, 3, srcs, self.wp[f"{nm}.{tag}.fwd"], P.p(f"{nm}.{tag}.bias"), [y], None, stats, self.conv_impl)
            rm, rv = self._buf_(f"{nm}.{bnt}.running_mean"), self._buf_(f"{nm}.{bnt}.running_var")
            if training:
                ops.bn_finalize(fl, count, stats, P.p(f"{nm}.{bnt}.weight"), P.p(f"{nm}.{bnt}.bias"), BN_EPS, BN_MOMENTUM,
                                rm, rv, sc, sh, mu, rs)
            else:
                g = P.p(f"{nm}.{bnt}.weight")
                torch.mul(g, torch.rsqrt(rv + BN_EPS), out=sc)
                torch.sub(P.p(f"{nm}.{bnt}.bias"), rm * sc, out=sh)
            if idx == 0:
                ops.bn_act(y, sc, sh, None, True, e.h, None)
            else:
                ops.bn_act(y, sc, sh, e.y1, True, e.out, e.pool)

    def _up_forward(self, l: int, j: int):
        nm = self.up_names[(l, j)]
        src = self.slot(self.X, l + 1, j)
        up = self.UP[(l, j)]
        h_, w_ = self._hw(l + 1)
        dsts = [up.phase(k // 2, k % 2) for k in range(4)]
        self.ops.conv2d(self.N, h_, w_, 1, [src], self.wp[f"{nm}.fwd"], self.wp[f"{nm}.bias4"], dsts, None, None, self.conv_impl)

    def forward(self, xA: torch.Tensor, xB: torch.Tensor, training: bool = True) -> torch.Tensor:
        ops, N, H, W, Cin = self.ops, self.N, self.H, self.W, self.in_ch
        assert tuple(xA.shape) == (N, Cin, H, W) and tuple(xB.shape) == (N, Cin, H, W), \
            f"engine was planned for {(N, Cin, H, W)}, got {tuple(xA.shape)}"
        self.params.ensure(self.device)
        for br, x in enumerate((xA, xB)):
            x = x.contiguous()
            if x.dtype != torch.float32:
                x = x.float()
            self._x_in[br] = x
            if not self.use_stem:
                ops.permute_cast(x, self.xin[br].base, (N, H, W, Cin), (Cin * H * W, W, 1, H * W))
        self._pack_weights()
        if training:
            ops.zero_(self.stats_all)
            self._ensure_nbt()
            self.nbt_all.add_(self.nbt_incr)      # one launch for the 30 num_batches_tracked counters
        for e in self.execs:
            if e.key[0] == "dec":
                self._up_forward(e.key[1], e.key[2])
            self._block_forward(e, training)
        xs = [self.slot(self.X, 0, 2 + i) for i in range(4)]
        P = self.params
        ops.ecam_pool(xs, self.pooled, self.argmax, self.pool_scratch)
        ops.ecam_gates(N, self.f[0], 4, self.hid, self.hid1, self.pooled, P.p("ca.fc1.weight"), P.p("ca.fc2.weight"),
                       P.p("ca1.fc1.weight"), P.p("ca1.fc2.weight"), self.gates, self.hidden)
        ops.ecam_final(xs, self.gates, P.p("conv_final.weight"), P.p("conv_final.bias"), self.K, self.logits,
                       self.pooled if training else None, self.argmax if training else None)
        return self.logits

    # ------------------------------------------------------------------------------------------
    def _block_backward(self, e: _Exec, gdsts: Optional[List[View]], gacc: Optional[List[bool]], first: bool,
                        dpool: Optional[View] = None):
        ops, P, N = self.ops, self.params, self.N
        h_, w_ = self._hw(e.level)
        fl = self.f[e.level]
        nm = e.name
        count = float(N * h_ * w_)
        sc1, sh1, mu1, rs1, sc2, sh2, mu2, rs2 = [e.bn[i * fl:(i + 1) * fl] for i in range(8)]
        dy, dh = self.dY[e.level], self.dH[e.level]
        acc = not first
        # ---- bn2 + residual + relu backward -> dy2.  Pass 1 masks dout IN PLACE (g = dout*(out>0)); the masked
        # gradient is re-used by pass 2 and, below, as the identity-path gradient of conv1's output.
        ops.bn_bwd_reduce(e.dout, e.out, e.y2, None, None, mu2, rs2, e.bstats[1], dpool)
        ops.bn_bwd_apply(e.dout, True, e.y2, None, None, mu2, rs2, P.p(f"{nm}.bn2.weight"), e.bstats[1], count, None, dy,
                         P.g(f"{nm}.bn2.weight"), P.g(f"{nm}.bn2.bias"), P.g(f"{nm}.conv1.bias"), acc)
        # d(conv2.bias) = sum(dy2) == 0 identically (BatchNorm removes the mean): the flat gradient buffer keeps its 0.
        ops.conv2d_wgrad(N, h_, w_, 3, [e.h], [dy], self.gp[f"{nm}.conv2"], acc, self.conv_impl)
        ops.conv2d(N, h_, w_, 3, [dy], self.wp[f"{nm}.conv2.dgrad"], None, [dh], None, None, self.conv_impl)
        # ---- bn1 + relu backward: the ReLU mask is recomputed from y1 (h is not re-read); + identity path g -> dy1
        ops.bn_bwd_reduce(dh, None, e.y1, sc1, sh1, mu1, rs1, e.bstats[0])
        ops.bn_bwd_apply(dh, False, e.y1, sc1, sh1, mu1, rs1, P.p(f"{nm}.bn1.weight"), e.bstats[0], count, e.dout, dy,
                         P.g(f"{nm}.bn1.weight"), P.g(f"{nm}.bn1.bias"), None, acc)
        if self.use_stem and e.key[0] == "enc" and e.level == 0:
            ops.stem_wgrad3x3(self._x_in[e.key[2]], dy, P.g(f"{nm}.conv1.weight"), acc)
        else:
            ops.conv2d_wgrad(N, h_, w_, 3, e.srcs, [dy], self.gp[f"{nm}.conv1"], acc, self.conv_impl)
        if gdsts:
            ops.conv2d(N, h_, w_, 3, [dy], self.wp[f"{nm}.conv1.dgrad"], None, gdsts, gacc, None, self.conv_impl)

    def backward(self, dlogits: torch.Tensor):
        ops, P, N, f = self.ops, self.params, self.N, self.f
        ops.zero_(self.bstats_all)
        xs = [self.slot(self.X, 0, 2 + i) for i in range(4)]
        dxs = [self.slot(self.dX, 0, 2 + i) for i in range(4)]
        ops.ecam_bwd_reduce(xs, self.K, dlogits, self.red)
        ops.ecam_gates_bwd(N, f[0], 4, self.hid, self.hid1, self.K, self.pooled, self.hidden, self.gates, self.red,
                           P.p("conv_final.weight"), P.p("ca.fc1.weight"), P.p("ca.fc2.weight"), P.p("ca1.fc1.weight"),
                           P.p("ca1.fc2.weight"), self.dpooled, P.g("conv_final.weight"), P.g("conv_final.bias"),
                           P.g("ca.fc1.weight"), P.g("ca.fc2.weight"), P.g("ca1.fc1.weight"), P.g("ca1.fc2.weight"), False)
        ops.ecam_bwd_apply(dxs, self.gates, P.p("conv_final.weight"), self.K, dlogits, self.dpooled, self.argmax)
        written = {(0, 2), (0, 3), (0, 4), (0, 5)}
        seen_blocks = set()
        for (l, j) in reversed(DEC_ORDER):
            e = self.exec_of[("dec", l, j)]
            # gradient destinations: runs of prefix slots with equal first-writer status, then the up tensor
            gd, ga = [], []
            k = 0
            while k <= j:
                st = (l, k) in written
                k2 = k
                while k2 + 1 <= j and ((l, k2 + 1) in written) == st:
                    k2 += 1
                if self.planar:
                    gd += [self.dX[l][kk] for kk in range(k, k2 + 1)]
                    ga += [st] * (k2 - k + 1)
                else:
                    gd.append(self.dX[l].ch(k * f[l], (k2 - k + 1) * f[l]))
                    ga.append(st)
                k = k2 + 1
            for kk in range(j + 1):
                written.add((l, kk))
            gd.append(self.dUP[(l, j)])
            ga.append(False)
            self._block_backward(e, gd, ga, e.name not in seen_blocks)
            seen_blocks.add(e.name)
            # transposed-conv backward
            nm = self.up_names[(l, j)]
            h1, w1 = self._hw(l + 1)
            dup = self.dUP[(l, j)]
            phases = [dup.phase(k_ // 2, k_ % 2) for k_ in range(4)]
            tgt = (l + 1, j)
            ops.conv2d(N, h1, w1, 1, phases, self.wp[f"{nm}.dgrad"], None, [self.slot(self.dX, l + 1, j)],
                       [tgt in written], None, self.conv_impl)
            written.add(tgt)
            # weight gradient of the four phases + the bias gradient (one bias per channel for all four phases: folded modulo C).  With
            # 64 input channels (the full-resolution level: 4 x 411 MB of dUP) the last 128-row tile of the weight-gradient GEMM has a free
            # group slot and the per-channel sums ride along in its UMMAs; wider layers keep the one dense pass over dUP.
            if f[l + 1] % 128 != 0 and self.dtype == torch.bfloat16:
                ops.conv2d_wgrad_bias(N, h1, w1, 1, [self.slot(self.X, l + 1, j)], phases, self.gp[nm], P.g(f"{nm}.up.bias"), f[l + 1],
                                      False, False, self.conv_impl)
            else:
                ops.conv2d_wgrad(N, h1, w1, 1, [self.slot(self.X, l + 1, j)], phases, self.gp[nm], False, self.conv_impl)
                ops.channel_sum(dup, P.g(f"{nm}.up.bias"), False)
        # Data-parallel overlap: everything registered from conv0_1 on (decoder blocks, their ConvTransposes, ECAM, classifier) is final
        # here, conv4_0 + Up4_0 (registered right before conv0_1) after the first encoder block below: the flat gradient range
        # [conv4_0 .., end) = 83 % of the bytes goes out under the rest of the encoder backward, the shared encoder blocks at the end.
        self._unpack_grads("dec")
        off = P.offsets
        if "conv0_1.conv1.weight" in off:
            self._grads_ready(off["conv0_1.conv1.weight"][0], P.numel)
        for br in (1, 0):
            for l in (4, 3, 2, 1, 0):
                if l == 4 and br == 0:
                    continue
                e = self.exec_of[("enc", l, br)]
                dpool = None
                if (l + 1, br) in self.P:
                    if (l, br) in written:
                        dpool = self.dP[(l + 1, br)]     # max-pool backward folded into the BN-backward reduce pass
                    else:
                        ops.maxpool2x2_bwd(e.out, self.dP[(l + 1, br)], e.dout, False)
                        written.add((l, br))
                assert (l, br) in written
                gd = None if l == 0 else [self.dP[(l, br)]]
                self._block_backward(e, gd, None if gd is None else [False], e.name not in seen_blocks, dpool)
                seen_blocks.add(e.name)
                if l == 4:
                    self._unpack_grads("l4")
                    if "conv4_0.conv1.weight" in off and "conv0_1.conv1.weight" in off:
                        self._grads_ready(off["conv4_0.conv1.weight"][0], off["conv0_1.conv1.weight"][0])
        self._unpack_grads("enc")
