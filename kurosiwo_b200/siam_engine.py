"""Host-side schedule of the FC-Siam-conc / FC-Siam-diff training step over the C-ABI ops.

Reference path being replaced: models/siam_conc.py:95-177, models/siam_diff.py:95-173 (forward) and their autograd
backward, inside training/change_detection_trainer.py:136-177.

Every layer of the reference is `Dropout2d(ReLU(BN(conv)))`:
  nn.Conv2d(k3,p1)                        -> ks_conv2d with [tap][Cout][Cin] weights
  nn.ConvTranspose2d(k3,s1,p1) (decoder)  -> the same conv with 180-degree-rotated, in/out-swapped weights
  nn.ConvTranspose2d(k3,s2,p1,op1)        -> ONE 3x3 conv launch producing 4*C channels, each group of C written to one 2x2
                                            output phase: out[2i+a] takes x[i] (a=0: ky=1; a=1: ky=2) and x[i+1] (a=1: ky=0);
                                            the unused (tap, phase) weight blocks stay zero
  torch.cat(up, skip_1, skip_2)           -> K-dimension view list (never materialised)
  |skip_1 - skip_2| (diff)                -> ks_absdiff_fwd / _bwd
  Softmax / LogSoftmax(dim=1)             -> ks_softmax_head_fwd / _bwd (NHWC storage dtype -> NCHW fp32, what the criterion reads)
  Dropout2d                               -> one ks_dropout_mask launch per step for all 29 executions, ks_channel_scale per tensor
Activations are NHWC in the storage dtype (bf16 perf mode / fp32 parity mode); gradients mirror them and producers
assign or accumulate according to a first-writer analysis done while the backward is issued.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .engine_common import FlatParams, TrainStepMixin
from .lib import View

ENC = [("11", 16), ("12", 16), ("21", 32), ("22", 32), ("31", 64), ("32", 64), ("33", 64), ("41", 128), ("42", 128), ("43", 128)]
STAGE_LAST = {"12": 0, "22": 1, "33": 2, "43": 3}          # layer -> pyramid level of the stage it closes
DEC = [("43d", 128, 3), ("42d", 128, 3), ("41d", 64, 3), ("33d", 64, 2), ("32d", 64, 2), ("31d", 32, 2),
       ("22d", 32, 1), ("21d", 16, 1), ("12d", 16, 0)]
UP_BEFORE = {"43d": "upconv4", "33d": "upconv3", "22d": "upconv2", "12d": "upconv1"}
SKIP_OF = {"43d": "43", "33d": "33", "22d": "22", "12d": "12"}
BN_EPS, BN_MOMENTUM = 1e-5, 0.1
HEAD_PAD = 16          # conv11d's output channels are padded to 16 (the conv engine's minimum N tile)
# (output phase a, input offset di, kernel index ky) per axis of ConvTranspose2d(k3,s2,p1,op1)
UP_TERMS = [(0, 0, 1), (1, 0, 2), (1, 1, 0)]


class _T:
    """An activation with its gradient mirror and the first-writer flag of the current backward."""
    __slots__ = ("v", "g", "written")

    def __init__(self, v: View, g: Optional[View]):
        self.v, self.g, self.written = v, g, False


class _Layer:
    __slots__ = ("name", "mask_key", "level", "cout", "srcs", "y", "out", "pool", "bn", "stats", "bstats", "transposed", "mask",
                 "pmask")


class SiamUNetEngine(TrainStepMixin):
    def __init__(self, ops, module: torch.nn.Module, kind: str, in_ch: int, num_classes: int, N: int, H: int, W: int,
                 dtype: torch.dtype, device, conv_impl: int = 0):
        assert kind in ("conc", "diff")
        assert H % 16 == 0 and W % 16 == 0, "the fused path needs H, W divisible by 16 (ReplicationPad2d is then zero-width)"
        assert num_classes == 3, "the fused head/loss kernels are built for num_classes == 3 (configs/config.json:13)"
        self.ops, self.module, self.kind, self.dtype, self.device = ops, module, kind, dtype, torch.device(device)
        self.in_ch, self.K, self.N, self.H, self.W = in_ch, num_classes, N, H, W
        self.conv_impl = conv_impl
        self.params = FlatParams(module)
        self.fixed_masks: Optional[Dict[str, torch.Tensor]] = None   # tests inject Dropout2d masks here
        self.dropout_seed = 0x5EED
        self._build()

    # ------------------------------------------------------------------------------------------
    def _hw(self, l):
        return self.H >> l, self.W >> l

    def _new(self, l, C, grad=True) -> _T:
        h, w = self._hw(l)
        v = View.alloc(self.N, h, w, C, self.dtype, self.device)
        g = View.alloc(self.N, h, w, C, self.dtype, self.device) if grad else None
        return _T(v, g)

    def _build(self):
        N, dev = self.N, self.device
        self.xin = [self._new(0, self.in_ch, grad=False) for _ in range(2)]
        self.layers: List[_Layer] = []
        self.skips: Dict = {}
        self.dy = {}                                    # per (level, C) scratch for the pre-BN gradient

        def layer(name, mask_key, level, cout, srcs, transposed, pool_level=None):
            L = _Layer()
            L.name, L.mask_key, L.level, L.cout, L.srcs, L.transposed = name, mask_key, level, cout, srcs, transposed
            L.y = self._new(level, cout, grad=False).v
            L.out = self._new(level, cout)
            L.pool = self._new(pool_level, cout) if pool_level is not None else None
            L.bn = torch.zeros(4 * cout, dtype=torch.float32, device=dev)      # scale, shift, mean, rstd
            if (level, cout) not in self.dy:
                self.dy[(level, cout)] = self._new(level, cout, grad=False).v
            self.layers.append(L)
            return L

        last_pool = None
        for br in (0, 1):                               # x1 then x2 (running-stat order, siam_conc.py:99-144)
            h = self.xin[br]
            level = 0
            for n, c in ENC:
                closes = n in STAGE_LAST
                want_pool = closes and not (n == "43" and br == 0)      # x4p_1 is computed and never used (:121)
                L = layer(n, f"{n}_{br + 1}", level, c, [h], False, level + 1 if want_pool else None)
                h = L.out
                if closes:
                    self.skips[(n, br)] = L.out
                    if want_pool:
                        h = L.pool
                    level += 1
            last_pool = h
        self.ups = []
        h = last_pool
        for n, c, level in DEC:
            srcs = [h]
            if n in UP_BEFORE:
                up = self._new(level, h.v.C)
                s1, s2 = self.skips[(SKIP_OF[n], 0)], self.skips[(SKIP_OF[n], 1)]
                diff = None
                if self.kind == "conc":
                    srcs = [up, s1, s2]
                else:
                    d = self._new(level, s1.v.C)
                    srcs, diff = [up, d], (s1, s2, d)
                self.ups.append((UP_BEFORE[n], h, up, level, diff))
            L = layer(n, n, level, c, srcs, True)
            h = L.out
        self.head_in = h
        self.z = self._new(0, HEAD_PAD)
        self.logits = torch.zeros(N, self.K, self.H, self.W, dtype=torch.float32, device=dev)   # model OUTPUT (probs / log-probs)
        self.up_of_layer = {}
        i = 0
        for L in self.layers:
            if L.name in UP_BEFORE:
                self.up_of_layer[L.name] = self.ups[i]
                i += 1
        # BatchNorm statistics workspaces (fp64 sums), flat so that one zero-fill clears them
        tot = sum(2 * L.cout for L in self.layers)
        self.stats_all = torch.zeros(tot, dtype=torch.float64, device=dev)
        self.bstats_all = torch.zeros(tot, dtype=torch.float64, device=dev)
        o = 0
        for L in self.layers:
            L.stats, L.bstats = self.stats_all[o:o + 2 * L.cout], self.bstats_all[o:o + 2 * L.cout]
            o += 2 * L.cout
        # Dropout2d masks: one flat fp32 buffer [sum_exec N*C]
        tot = sum(N * L.cout for L in self.layers)
        self.masks_all = torch.ones(tot, dtype=torch.float32, device=dev)
        o = 0
        for L in self.layers:
            L.mask = self.masks_all[o:o + N * L.cout]
            o += N * L.cout
        self.do_step = torch.zeros(1, dtype=torch.int32, device=dev)
        # packed weights (storage dtype) / packed weight gradients (fp32), zero-initialised: unused blocks stay zero
        self.wp: Dict[str, torch.Tensor] = {}
        self.gp: Dict[str, torch.Tensor] = {}
        self.cin_of = {}
        for L in self.layers:
            if L.name in self.cin_of:
                continue
            cin = sum(s.v.C for s in L.srcs)
            self.cin_of[L.name] = cin
            for tag in ("fwd", "dgrad"):
                self.wp[f"conv{L.name}.{tag}"] = torch.zeros(9 * L.cout * cin, dtype=self.dtype, device=dev)
            self.gp[f"conv{L.name}"] = torch.zeros(9 * L.cout * cin, dtype=torch.float32, device=dev)
        for (nm, src, up, level, _) in self.ups:
            c = src.v.C
            self.wp[f"{nm}.fwd"] = torch.zeros(9 * 4 * c * c, dtype=self.dtype, device=dev)
            self.wp[f"{nm}.dgrad"] = torch.zeros(9 * c * 4 * c, dtype=self.dtype, device=dev)
            self.wp[f"{nm}.bias4"] = torch.zeros(4 * c, dtype=torch.float32, device=dev)
            self.gp[nm] = torch.zeros(9 * 4 * c * c, dtype=torch.float32, device=dev)
        ch = self.head_in.v.C
        self.wp["conv11d.fwd"] = torch.zeros(9 * HEAD_PAD * ch, dtype=self.dtype, device=dev)
        self.wp["conv11d.dgrad"] = torch.zeros(9 * ch * HEAD_PAD, dtype=self.dtype, device=dev)
        self.wp["conv11d.bias"] = torch.zeros(HEAD_PAD, dtype=torch.float32, device=dev)
        self.gp["conv11d"] = torch.zeros(9 * HEAD_PAD * ch, dtype=torch.float32, device=dev)
        self.gp["conv11d.bias"] = torch.zeros(HEAD_PAD, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------------------------------
    def _pack_jobs(self):
        P, jobs = self.params, []
        for L in self.layers:
            nm, co, ci = f"conv{L.name}", L.cout, self.cin_of[L.name]
            if any(j[1] is self.wp[f"{nm}.fwd"] for j in jobs):
                continue                                               # shared encoder layer: pack once
            w = P.p(f"{nm}.weight")
            if not L.transposed:      # nn.Conv2d (O,I,3,3)
                jobs.append((w, self.wp[f"{nm}.fwd"], (9, co, ci), (1, ci * 9, 9), 0))            # [t][o][i] = w[o][i][t]
                jobs.append((w, self.wp[f"{nm}.dgrad"], (9, ci, co), (-1, 9, ci * 9), 8))         # [t][i][o] = w[o][i][8-t]
            else:                     # nn.ConvTranspose2d (I,O,3,3), k3 s1 p1 == conv with w_c[o][i][t] = w[i][o][8-t]
                jobs.append((w, self.wp[f"{nm}.fwd"], (9, co, ci), (-1, 9, co * 9), 8))           # [t][o][i] = w[i][o][8-t]
                jobs.append((w, self.wp[f"{nm}.dgrad"], (9, ci, co), (1, co * 9, 9), 0))          # [t][i][o] = w[i][o][t]
        for (nm, src, up, level, _) in self.ups:
            c = src.v.C
            w = P.p(f"{nm}.weight")   # (Cin, Cout, 3, 3)
            for (a, di, ky) in UP_TERMS:
                for (b, dj, kx) in UP_TERMS:
                    ph, off = a * 2 + b, ky * 3 + kx
                    tap = (di + 1) * 3 + (dj + 1)          # forward: out_phase[i,j] += x[i+di, j+dj] . w[:, :, ky, kx]
                    jobs.append((w, self.wp[f"{nm}.fwd"], (c, c), (9, c * 9), off, (c, 1), (tap * 4 * c + ph * c) * c))
                    tapd = (1 - di) * 3 + (1 - dj)         # data gradient: dx[i,j] += dup_phase[i-di, j-dj] . w^T
                    jobs.append((w, self.wp[f"{nm}.dgrad"], (c, c), (c * 9, 9), off, (4 * c, 1), tapd * c * 4 * c + ph * c))
            jobs.append((P.p(f"{nm}.bias"), self.wp[f"{nm}.bias4"], (4, c), (0, 1), 0))
        ch, K = self.head_in.v.C, self.K
        w = P.p("conv11d.weight")     # ConvTranspose2d (16, K, 3, 3), output channels padded to HEAD_PAD
        jobs.append((w, self.wp["conv11d.fwd"], (9, K, ch), (-1, 9, K * 9), 8, (HEAD_PAD * ch, ch, 1), 0))
        jobs.append((w, self.wp["conv11d.dgrad"], (9, ch, K), (1, K * 9, 9), 0, (ch * HEAD_PAD, HEAD_PAD, 1), 0))
        jobs.append((P.p("conv11d.bias"), self.wp["conv11d.bias"], (K,), (1,), 0))
        return jobs

    def _unpack_jobs(self):
        P, jobs, seen = self.params, [], set()
        for L in self.layers:
            nm, co, ci = f"conv{L.name}", L.cout, self.cin_of[L.name]
            if nm in seen:
                continue
            seen.add(nm)
            if not L.transposed:      # grad[o][i][t] = gp[t][o][i]
                jobs.append((self.gp[nm], P.g(f"{nm}.weight"), (co, ci, 9), (ci, 1, co * ci), 0))
            else:                     # grad[i][o][t] = gp[8-t][o][i]
                jobs.append((self.gp[nm], P.g(f"{nm}.weight"), (ci, co, 9), (1, ci, -co * ci), 8 * co * ci))
        for (nm, src, up, level, _) in self.ups:
            c = src.v.C
            for (a, di, ky) in UP_TERMS:
                for (b, dj, kx) in UP_TERMS:
                    ph, off = a * 2 + b, ky * 3 + kx
                    tap = (di + 1) * 3 + (dj + 1)          # grad[ci][co][ky][kx] = gp[tap][ph*C+co][ci]
                    jobs.append((self.gp[nm], P.g(f"{nm}.weight"), (c, c), (1, c), (tap * 4 * c + ph * c) * c, (c * 9, 9), off))
        ch, K = self.head_in.v.C, self.K
        jobs.append((self.gp["conv11d"], P.g("conv11d.weight"), (ch, K, 9), (1, ch, -HEAD_PAD * ch), 8 * HEAD_PAD * ch))
        jobs.append((self.gp["conv11d.bias"], P.g("conv11d.bias"), (K,), (1,), 0))
        return jobs

    def _tables(self):
        key = (self.params.flat.data_ptr(), self.params.grad.data_ptr())
        if getattr(self, "_table_key", None) != key:
            self._pack_table = self.ops.make_permute_table(self._pack_jobs(), self.device)
            self._unpack_table = self.ops.make_permute_table(self._unpack_jobs(), self.device)
            self._table_key = key
        return self._pack_table, self._unpack_table

    def _bn_modules(self):
        seen, mods = set(), []
        for L in self.layers:
            if L.name not in seen:
                seen.add(L.name)
                mods.append((L.name, self.module.get_submodule(f"bn{L.name}")))
        return mods

    def _ensure_nbt(self):
        mods = self._bn_modules()
        flat = getattr(self, "nbt_all", None)
        ok = flat is not None and flat.device == self.device and all(
            m.num_batches_tracked.data_ptr() == flat.data_ptr() + 8 * i for i, (_, m) in enumerate(mods))
        if ok:
            return
        flat = torch.zeros(len(mods), dtype=torch.int64, device=self.device)
        incr = torch.zeros(len(mods), dtype=torch.int64, device=self.device)
        for i, (name, m) in enumerate(mods):
            flat[i] = m.num_batches_tracked.to(self.device)
            m._buffers["num_batches_tracked"] = flat[i]
            incr[i] = sum(1 for L in self.layers if L.name == name)
        self.nbt_all, self.nbt_incr = flat, incr

    # ------------------------------------------------------------------------------------------
    def _dropout_p(self) -> float:
        return float(getattr(self.module, "dropout_p", 0.2))

    def _prepare_masks(self, training: bool) -> bool:
        """Returns True when Dropout2d is active for this forward."""
        if not training:
            return False
        if self.fixed_masks is not None:
            for L in self.layers:
                L.mask.copy_(self.fixed_masks[L.mask_key].reshape(-1).to(self.device, torch.float32))
            return True
        p = self._dropout_p()
        if p <= 0.0:
            return False
        self.do_step += 1
        self.ops.dropout_mask(self.masks_all, p, self.dropout_seed, self.do_step)
        return True

    def forward(self, x1: torch.Tensor, x2: torch.Tensor, training: bool = True) -> torch.Tensor:
        ops, N, H, W, Cin, P = self.ops, self.N, self.H, self.W, self.in_ch, self.params
        assert tuple(x1.shape) == (N, Cin, H, W) and tuple(x2.shape) == (N, Cin, H, W), \
            f"engine was planned for {(N, Cin, H, W)}, got {tuple(x1.shape)}"
        P.ensure(self.device)
        for br, x in enumerate((x1, x2)):
            x = x.contiguous()
            if x.dtype != torch.float32:
                x = x.float()
            ops.permute_cast(x, self.xin[br].v.base, (N, H, W, Cin), (Cin * H * W, W, 1, H * W))
        ops.permute_cast_table(self._tables()[0])
        self._dropping = self._prepare_masks(training)
        if training:
            ops.zero_(self.stats_all)
            self._ensure_nbt()
            self.nbt_all.add_(self.nbt_incr)
        for L in self.layers:
            if L.name in UP_BEFORE:
                self._up_forward(self.up_of_layer[L.name])
            self._layer_forward(L, training)
        ch = self.head_in.v.C
        ops.conv2d(N, H, W, 3, [self.head_in.v], self.wp["conv11d.fwd"], self.wp["conv11d.bias"], [self.z.v], None, None, self.conv_impl)
        ops.softmax_head_fwd(self.z.v, self.K, self.kind == "diff", self.logits)
        return self.logits

    def _up_forward(self, u):
        nm, src, up, level, diff = u
        h_, w_ = self._hw(level + 1)
        dsts = [up.v.phase(k // 2, k % 2) for k in range(4)]
        self.ops.conv2d(self.N, h_, w_, 3, [src.v], self.wp[f"{nm}.fwd"], self.wp[f"{nm}.bias4"], dsts, None, None, self.conv_impl)
        if diff is not None:
            s1, s2, d = diff
            self.ops.absdiff_fwd(s1.v, s2.v, d.v)

    def _layer_forward(self, L: _Layer, training: bool):
        ops, P, N = self.ops, self.params, self.N
        h_, w_ = self._hw(L.level)
        c, nm = L.cout, L.name
        sc, sh, mu, rs = [L.bn[i * c:(i + 1) * c] for i in range(4)]
        stats = L.stats if training else None
        ops.conv2d(N, h_, w_, 3, [s.v for s in L.srcs], self.wp[f"conv{nm}.fwd"], P.p(f"conv{nm}.bias"), [L.y], None, stats, self.conv_impl)
        bnm = self.module.get_submodule(f"bn{nm}")
        if training:
            ops.bn_finalize(c, float(N * h_ * w_), stats, P.p(f"bn{nm}.weight"), P.p(f"bn{nm}.bias"), BN_EPS, BN_MOMENTUM,
                            bnm.running_mean, bnm.running_var, sc, sh, mu, rs)
        else:
            torch.mul(P.p(f"bn{nm}.weight"), torch.rsqrt(bnm.running_var + BN_EPS), out=sc)
            torch.sub(P.p(f"bn{nm}.bias"), bnm.running_mean * sc, out=sh)
        ops.bn_act(L.y, sc, sh, None, True, L.out.v, L.pool.v if L.pool is not None else None)
        if self._dropping:
            ops.channel_scale(L.out.v, L.mask)
            if L.pool is not None:
                ops.channel_scale(L.pool.v, L.mask)       # maxpool(m*x) == m*maxpool(x) for the non-negative mask

    # ------------------------------------------------------------------------------------------
    def _grad_dsts(self, srcs: List[_T]):
        gd, ga = [], []
        for s in srcs:
            gd.append(s.g)
            ga.append(s.written)
            s.written = True
        return gd, ga

    def _layer_backward(self, L: _Layer, first: bool):
        ops, P, N = self.ops, self.params, self.N
        h_, w_ = self._hw(L.level)
        c, nm = L.cout, L.name
        sc, sh, mu, rs = [L.bn[i * c:(i + 1) * c] for i in range(4)]
        dy = self.dy[(L.level, c)]
        acc = not first
        dout = L.out.g
        dpool = None
        if L.pool is not None and L.pool.written:
            if not L.out.written or self._dropping:
                ops.maxpool2x2_bwd(L.out.v, L.pool.g, dout, L.out.written)
                L.out.written = True
            else:
                dpool = L.pool.g                       # folded into the BN-backward reduce pass
        assert L.out.written, f"no gradient reached layer {nm}"
        if self._dropping:
            ops.channel_scale(dout, L.mask)
        ops.bn_bwd_reduce(dout, L.out.v, L.y, None, None, mu, rs, L.bstats, dpool)
        ops.bn_bwd_apply(dout, True, L.y, None, None, mu, rs, P.p(f"bn{nm}.weight"), L.bstats, float(N * h_ * w_), None, dy,
                         P.g(f"bn{nm}.weight"), P.g(f"bn{nm}.bias"), None, acc)
        # d(conv.bias) = sum(dy) == 0 identically (the BatchNorm removes the mean): the flat gradient keeps its 0.
        ops.conv2d_wgrad(N, h_, w_, 3, [s.v for s in L.srcs], [dy], self.gp[f"conv{nm}"], acc, self.conv_impl)
        if L.srcs[0].g is not None:
            gd, ga = self._grad_dsts(L.srcs)
            ops.conv2d(N, h_, w_, 3, [dy], self.wp[f"conv{nm}.dgrad"], None, gd, ga, None, self.conv_impl)

    def _up_backward(self, u):
        ops, P, N = self.ops, self.params, self.N
        nm, src, up, level, diff = u
        h_, w_ = self._hw(level + 1)
        if diff is not None:
            s1, s2, d = diff
            ops.absdiff_bwd(s1.v, s2.v, d.g, s1.g, s1.written, s2.g, s2.written)
            s1.written = s2.written = True
        phases = [up.g.phase(k // 2, k % 2) for k in range(4)]
        ops.conv2d(N, h_, w_, 3, phases, self.wp[f"{nm}.dgrad"], None, [src.g], [src.written], None, self.conv_impl)
        src.written = True
        ops.conv2d_wgrad(N, h_, w_, 3, [src.v], phases, self.gp[nm], False, self.conv_impl)
        ops.channel_sum(up.g, P.g(f"{nm}.bias"), False)

    def backward(self, dout: torch.Tensor):
        ops, N, H, W = self.ops, self.N, self.H, self.W
        ops.zero_(self.bstats_all)
        for L in self.layers:
            L.out.written = False
            if L.pool is not None:
                L.pool.written = False
        for u in self.ups:
            u[2].written = False
            if u[4] is not None:
                u[4][2].written = False
        ops.softmax_head_bwd(self.logits, dout, self.K, self.kind == "diff", self.z.g)
        ops.conv2d_wgrad(N, H, W, 3, [self.head_in.v], [self.z.g], self.gp["conv11d"], False, self.conv_impl)
        ops.channel_sum(self.z.g, self.gp["conv11d.bias"], False)
        ops.conv2d(N, H, W, 3, [self.z.g], self.wp["conv11d.dgrad"], None, [self.head_in.g], [False], None, self.conv_impl)
        self.head_in.written = True
        seen = set()
        for L in reversed(self.layers):
            self._layer_backward(L, L.name not in seen)
            seen.add(L.name)
            if L.name in UP_BEFORE:
                self._up_backward(self.up_of_layer[L.name])
        ops.permute_cast_table(self._tables()[1])
