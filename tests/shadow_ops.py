"""Reference semantics of every C-ABI op, written with torch CPU ops.  TEST-ONLY.

Two uses: (1) `-m "not gpu"` tests drive the product's host schedule (snunet_engine.py) through
this backend to check the wiring against the oracle without a GPU; (2) `-m gpu` tests compare
each CUDA kernel with the same contract on identical inputs.  Never imported by the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from kurosiwo_b200.lib import View


def _t(v: View) -> torch.Tensor:
    return v.tensor()


class ShadowOps:
    name = "shadow"

    def __init__(self):
        self.launches = 0

    def set_option(self, name, value):
        pass

    def zero_(self, t):
        t.zero_()

    # -- plumbing ----------------------------------------------------------------------------
    def permute_cast(self, src, dst, dims, strides, accumulate=False, src_offset=0):
        d = list(dims) + [1] * (4 - len(dims))
        s = list(strides) + [0] * (4 - len(strides))
        idx = torch.zeros(d, dtype=torch.int64)
        for ax in range(4):
            shape = [1, 1, 1, 1]
            shape[ax] = d[ax]
            idx = idx + (torch.arange(d[ax]) * s[ax]).view(shape)
        vals = src.reshape(-1)[(idx + src_offset).reshape(-1)].float()
        out = dst.reshape(-1)[: vals.numel()]
        if accumulate:
            vals = vals + out.float()
        out.copy_(vals.to(dst.dtype))

    def make_permute_table(self, jobs, device):
        return list(jobs)

    def permute_cast_table(self, table):
        for job in table:
            src, dst, dims, strides, off = job[:5]
            if len(job) > 7 and job[7] is not None:          # value multiplier
                src = src.float() * float(job[7])
            if len(job) > 5 and job[5] is not None:
                d = list(dims) + [1] * (4 - len(dims))
                tmp = torch.zeros(d[0] * d[1] * d[2] * d[3], dtype=torch.float32)
                self.permute_cast(src, tmp, dims, strides, False, off)
                ds = list(job[5]) + [0] * (4 - len(job[5]))
                doff = job[6] if len(job) > 6 else 0
                view = torch.as_strided(dst.reshape(-1), d, ds, dst.storage_offset() + doff)
                view.copy_(tmp.view(d).to(dst.dtype))
            else:
                self.permute_cast(src, dst, dims, strides, False, off)

    # -- convolution -------------------------------------------------------------------------
    def conv2d(self, N, H, W, ksize, srcs, weight, bias, dsts, acc=None, stats=None, impl=0):
        acc = acc or [False] * len(dsts)
        x = torch.cat([_t(v).float() for v in srcs], dim=3).permute(0, 3, 1, 2)
        cin = x.shape[1]
        cout = sum(v.C for v in dsts)
        w = weight.float().view(ksize * ksize, cout, cin).permute(1, 2, 0).reshape(cout, cin, ksize, ksize)
        y = F.conv2d(x, w, None if bias is None else bias.float(), padding=ksize // 2).permute(0, 2, 3, 1)
        c0 = 0
        for v, a in zip(dsts, acc):
            part = y[..., c0:c0 + v.C]
            t = _t(v)
            if a:
                part = part + t.float()
            t.copy_(part.to(t.dtype))
            c0 += v.C
        if stats is not None:
            st = torch.cat([_t(v).float() for v in dsts], dim=3).double().reshape(-1, cout)
            stats.view(2, cout)[0] += st.sum(0)
            stats.view(2, cout)[1] += (st * st).sum(0)

    def conv2d_wgrad(self, N, H, W, ksize, xs, dys, dw, accumulate=False, impl=0):
        x = torch.cat([_t(v).float() for v in xs], dim=3).permute(0, 3, 1, 2)
        dy = torch.cat([_t(v).float() for v in dys], dim=3).permute(0, 3, 1, 2)
        cin, cout = x.shape[1], dy.shape[1]
        g = torch.nn.grad.conv2d_weight(x, (cout, cin, ksize, ksize), dy, padding=ksize // 2)  # (co, ci, k, k)
        g = g.reshape(cout, cin, ksize * ksize).permute(2, 0, 1).reshape(-1)
        out = dw.reshape(-1)[: g.numel()]
        out.copy_(g + out if accumulate else g)

    def conv2d_wgrad_bias(self, N, H, W, ksize, xs, dys, dw, dbias, bias_mod=0, accumulate=False, accumulate_bias=False, impl=0):
        self.conv2d_wgrad(N, H, W, ksize, xs, dys, dw, accumulate, impl)
        cout = sum(v.C for v in dys)
        blen = bias_mod if bias_mod > 0 else cout
        out = dbias.reshape(-1)[:blen]
        if not accumulate_bias:
            out.zero_()
        c0 = 0
        for v in dys:
            o = c0 % bias_mod if bias_mod > 0 else c0
            out[o:o + v.C] += _t(v).float().reshape(-1, v.C).sum(0)
            c0 += v.C

    def stem_conv3x3(self, x_nchw, w_oihw, bias, dst, stats=None):
        t = _t(dst)
        q = (lambda a: a.to(t.dtype).float())
        y = F.conv2d(q(x_nchw.float()), q(w_oihw.view(dst.C, x_nchw.shape[1], 3, 3)), None if bias is None else bias.float(), padding=1)
        t.copy_(y.permute(0, 2, 3, 1).to(t.dtype))
        if stats is not None:
            self.bn_stats(dst, stats)

    def stem_wgrad3x3(self, x_nchw, dy, dw_oihw, accumulate=False):
        t = _t(dy)
        x = x_nchw.float().to(t.dtype).float()
        g = torch.nn.grad.conv2d_weight(x, (dy.C, x.shape[1], 3, 3), t.float().permute(0, 3, 1, 2), padding=1).reshape(-1)
        out = dw_oihw.reshape(-1)[: g.numel()]
        out.copy_(g + out if accumulate else g)

    # -- batch norm / pooling ----------------------------------------------------------------
    def bn_stats(self, x, sums):
        t = _t(x).double().reshape(-1, x.C)
        sums.view(2, x.C)[0] += t.sum(0)
        sums.view(2, x.C)[1] += (t * t).sum(0)

    def bn_finalize(self, Cn, count, sums, gamma, beta, eps, momentum, rmean, rvar, scale, shift, mean, rstd):
        s = sums.view(2, Cn)
        mu = s[0] / count
        var = (s[1] / count - mu * mu).clamp_min(0)
        rs = (1.0 / torch.sqrt(var + eps)).float()
        sc = gamma * rs
        scale.copy_(sc)
        shift.copy_(beta - mu.float() * sc)
        mean.copy_(mu.float())
        rstd.copy_(rs)
        if rmean is not None:
            rmean.mul_(1 - momentum).add_(momentum * mu.float())
        if rvar is not None:
            unb = var * count / (count - 1.0) if count > 1 else var
            rvar.mul_(1 - momentum).add_(momentum * unb.float())

    def bn_act(self, y, scale, shift, res, relu, out, pool):
        v = _t(y).float() * scale + shift
        if res is not None:
            v = v + _t(res).float()
        if relu:
            v = F.relu(v)
        o = _t(out)
        o.copy_(v.to(o.dtype))
        if pool is not None:
            p = F.max_pool2d(o.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
            _t(pool).copy_(p.to(o.dtype))

    def bn_bwd_reduce(self, dout, out, y, scale, shift, mean, rstd, sums, dpool=None):
        if dpool is not None:
            self.maxpool2x2_bwd(out, dpool, dout, True)
        yt = _t(y).float()
        if out is not None:
            mask = _t(out).float() > 0
        else:
            mask = torch.addcmul(shift, yt, scale) > 0
        g = _t(dout).float() * mask
        if out is not None:
            d = _t(dout)
            d.copy_(g.to(d.dtype))
        xh = (yt - mean) * rstd
        Cn = y.C
        sums.view(2, Cn)[0] += g.double().reshape(-1, Cn).sum(0)
        sums.view(2, Cn)[1] += (g * xh).double().reshape(-1, Cn).sum(0)

    def bn_bwd_apply(self, g, premasked, y, scale, shift, mean, rstd, gamma, sums, count, add, dy, dgamma, dbeta, dsum_out, accumulate):
        Cn = y.C
        s = sums.view(2, Cn).float()
        yt = _t(y).float()
        gg = _t(g).float()
        if not premasked:
            gg = gg * (torch.addcmul(shift, yt, scale) > 0)
        xh = (yt - mean) * rstd
        v = gamma * rstd * (gg - s[0] / count - xh * (s[1] / count))
        if add is not None:
            v = v + _t(add).float()
        t = _t(dy)
        t.copy_(v.to(t.dtype))
        for dst, val in ((dgamma, s[1]), (dbeta, s[0]), (dsum_out, s[0])):
            if dst is not None:
                dst.copy_(dst + val if accumulate else val)

    def maxpool2x2_bwd(self, x, dpool, dx, accumulate):
        xt = _t(x).float().permute(0, 3, 1, 2)
        _, idx = F.max_pool2d(xt, 2, 2, return_indices=True)
        g = F.max_unpool2d(_t(dpool).float().permute(0, 3, 1, 2), idx, 2, 2, output_size=xt.shape[-2:]).permute(0, 2, 3, 1)
        t = _t(dx)
        t.copy_((g + t.float() if accumulate else g).to(t.dtype))

    def channel_sum(self, x, out, accumulate):
        s = _t(x).float().reshape(-1, x.C).sum(0)
        out.copy_(out + s if accumulate else s)

    # -- ECAM --------------------------------------------------------------------------------
    def ecam_pool(self, xs, pooled, argmax, scratch):
        N, J, Cb = xs[0].N, len(xs), xs[0].C
        CT = (J + 1) * Cb
        ts = [_t(v).float() for v in xs]
        full = torch.cat(ts + [sum(ts)], dim=3).reshape(N, -1, CT)  # [N, HW, CT]
        p = pooled.view(N, 2, CT)
        p[:, 0] = full.mean(1)
        mx, am = full.max(1)
        # first occurrence of the maximum
        first = (full == mx.unsqueeze(1)).float().argmax(1)
        p[:, 1] = mx
        argmax.view(N, CT).copy_(first.to(torch.int32))

    def ecam_gates(self, N, Cb, J, hid, hid1, pooled, w_fc1, w_fc2, w1_fc1, w1_fc2, gates, hidden):
        CC, CT, HT = J * Cb, (J + 1) * Cb, hid + hid1
        p = pooled.view(N, 2, CT)
        g = gates.view(N, CT)
        hd = hidden.view(N, 2, HT)
        for (c0, c1, h0, h1, w1, w2) in ((0, CC, 0, hid, w_fc1.view(hid, CC), w_fc2.view(CC, hid)),
                                         (CC, CT, hid, HT, w1_fc1.view(hid1, Cb), w1_fc2.view(Cb, hid1))):
            ha = p[:, 0, c0:c1] @ w1.t()
            hm = p[:, 1, c0:c1] @ w1.t()
            hd[:, 0, h0:h1] = ha
            hd[:, 1, h0:h1] = hm
            g[:, c0:c1] = torch.sigmoid((F.relu(ha) + F.relu(hm)) @ w2.t())

    def ecam_final(self, xs, gates, wf, bf, K, logits, pooled=None, argmax=None):
        N, J, Cb = xs[0].N, len(xs), xs[0].C
        CC = J * Cb
        g = gates.view(N, (J + 1) * Cb)
        cat = torch.cat([_t(v).float() for v in xs], dim=3)  # N,H,W,CC
        out = g[:, None, None, :CC] * (cat + g[:, None, None, CC:].repeat(1, 1, 1, J))
        lg = out @ wf.view(K, CC).t() + bf
        logits.copy_(lg.permute(0, 3, 1, 2))

    def ecam_bwd_reduce(self, xs, K, dlogits, red):
        N, J, Cb = xs[0].N, len(xs), xs[0].C
        CC = J * Cb
        cat = torch.cat([_t(v).float() for v in xs], dim=3).reshape(N, -1, CC).double()
        dl = dlogits.reshape(N, K, -1).double()
        r = red.view(N, K * CC + K)
        r[:, :K * CC] = torch.bmm(dl, cat).reshape(N, K * CC)
        r[:, K * CC:] = dl.sum(2)

    def ecam_gates_bwd(self, N, Cb, J, hid, hid1, K, pooled, hidden, gates, red, wf, w_fc1, w_fc2, w1_fc1, w1_fc2,
                       dpooled, dwf, dbf, dw_fc1, dw_fc2, dw1_fc1, dw1_fc2, accumulate=False):
        CC, CT, HT = J * Cb, (J + 1) * Cb, hid + hid1
        p, hd, g = pooled.view(N, 2, CT), hidden.view(N, 2, HT), gates.view(N, CT)
        r = red.view(N, K * CC + K).float()
        B, D = r[:, :K * CC].view(N, K, CC), r[:, K * CC:]
        wfm = wf.view(K, CC)
        ca, ca1 = g[:, :CC], g[:, CC:]
        A = B + ca1.repeat(1, J)[:, None, :] * D[:, :, None]
        dca = (wfm[None] * A).sum(1)
        dca1 = (ca * (D @ wfm)).view(N, J, Cb).sum(1)
        outs = {"dwf": (ca[:, None, :] * A).sum(0).reshape(-1), "dbf": D.sum(0)}
        dp = dpooled.view(N, 2, CT)
        for (c0, c1, h0, h1, w1, w2, dg, gg, n1, n2) in (
                (0, CC, 0, hid, w_fc1.view(hid, CC), w_fc2.view(CC, hid), dca, ca, "dw_fc1", "dw_fc2"),
                (CC, CT, hid, HT, w1_fc1.view(hid1, Cb), w1_fc2.view(Cb, hid1), dca1, ca1, "dw1_fc1", "dw1_fc2")):
            ds = dg * gg * (1 - gg)
            ha, hm = hd[:, 0, h0:h1], hd[:, 1, h0:h1]
            outs[n2] = (ds.t() @ (F.relu(ha) + F.relu(hm))).reshape(-1)
            drelu = ds @ w2
            dha, dhm = drelu * (ha > 0), drelu * (hm > 0)
            outs[n1] = (dha.t() @ p[:, 0, c0:c1] + dhm.t() @ p[:, 1, c0:c1]).reshape(-1)
            dp[:, 0, c0:c1] = dha @ w1
            dp[:, 1, c0:c1] = dhm @ w1
        for name, dst in (("dwf", dwf), ("dbf", dbf), ("dw_fc1", dw_fc1), ("dw_fc2", dw_fc2), ("dw1_fc1", dw1_fc1), ("dw1_fc2", dw1_fc2)):
            if dst is not None:
                d = dst.reshape(-1)
                d.copy_(d + outs[name] if accumulate else outs[name])

    def ecam_bwd_apply(self, dxs, gates, wf, K, dlogits, dpooled, argmax):
        N, J, Cb = dxs[0].N, len(dxs), dxs[0].C
        H, W = dxs[0].H, dxs[0].W
        CC, CT = J * Cb, (J + 1) * Cb
        g = gates.view(N, CT)
        dp = dpooled.view(N, 2, CT)
        am = argmax.view(N, CT).long()
        dl = dlogits.reshape(N, K, H * W)
        dO = torch.einsum("kc,nkp->npc", wf.view(K, CC), dl)  # N,HW,CC
        v = g[:, None, :CC] * dO + (dp[:, 0, :CC] + dp[:, 0, CC:].repeat(1, J))[:, None, :] / (H * W)
        onehot = torch.zeros(N, H * W, CT)
        onehot.scatter_(1, am[:, None, :], 1.0)
        v = v + onehot[:, :, :CC] * dp[:, 1, None, :CC] + onehot[:, :, CC:].repeat(1, 1, J) * dp[:, 1, None, CC:].repeat(1, 1, J)
        v = v.view(N, H, W, CC)
        for j, d in enumerate(dxs):
            t = _t(d)
            t.copy_(v[..., j * Cb:(j + 1) * Cb].to(t.dtype))

    # -- loss / optimizer --------------------------------------------------------------------
    def ce_dice_workspace(self, N, device):
        return torch.empty(2 * N + 16, dtype=torch.float64, device=device)

    def ce_dice(self, logits, labels, class_weights, ignore_index, grad_scale, loss_out, dlogits, pred, workspace, dice_weight=1.0):
        from oracle.loss_oracle import ce_dice
        r = ce_dice(logits.numpy(), labels.numpy(), class_weights.numpy(), ignore_index)
        if dice_weight != 1.0:       # loss = dice_weight*Dice + CE: the CE part alone comes from torch's own cross_entropy
            z = logits.detach().clone().requires_grad_(True)
            ce = F.cross_entropy(z, labels, weight=class_weights, ignore_index=ignore_index)
            ce.backward()
            d_only = r["dlogits"] - z.grad.numpy()
            r = dict(r, loss=dice_weight * r["dice"] + r["ce"], dlogits=dice_weight * d_only + z.grad.numpy())
        loss_out.copy_(torch.tensor([r["loss"], r["dice"], r["ce"]], dtype=torch.float32))
        if dlogits is not None:
            dlogits.copy_(torch.from_numpy(r["dlogits"] * grad_scale).float())
        if pred is not None:
            pred.copy_(torch.from_numpy(r["argmax"]))

    def adam_step(self, p, g, m, v, lr, b1, b2, eps, wd, grad_scale, step):
        t = int(step.item()) + 1
        gg = g * grad_scale + wd * p
        m.mul_(b1).add_(gg, alpha=1 - b1)
        v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        p.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)
        step += 1

    def adamw_step(self, p, g, m, v, lr, b1, b2, eps, wd, grad_scale, step):
        p.mul_(1 - lr * wd)                                   # torch.optim.AdamW: decoupled decay first
        self.adam_step(p, g, m, v, lr, b1, b2, eps, 0.0, grad_scale, step)

    # -- Siamese U-Net passes ------------------------------------------------------------------
    def softmax_head_fwd(self, z, K, log_mode, out):
        zz = _t(z).float()[..., :K].permute(0, 3, 1, 2)
        out.copy_(F.log_softmax(zz, 1) if log_mode else F.softmax(zz, 1))

    def softmax_head_bwd(self, out, dout, K, log_mode, dz):
        if log_mode:
            g = dout - out.exp() * dout.sum(1, keepdim=True)
        else:
            g = out * (dout - (dout * out).sum(1, keepdim=True))
        t = _t(dz)
        t.zero_()
        t[..., :K].copy_(g.permute(0, 2, 3, 1).to(t.dtype))

    def dropout_mask(self, mask, p, seed, step):
        gen = torch.Generator().manual_seed(int(seed) + (0 if step is None else int(step.item())))
        mask.copy_((torch.rand(mask.numel(), generator=gen) >= p).float() / (1.0 - p))

    def channel_scale(self, x, m):
        t = _t(x)
        t.copy_((t.float() * m.view(x.N, 1, 1, x.C)).to(t.dtype))

    def absdiff_fwd(self, a, b, out):
        t = _t(out)
        t.copy_((_t(a).float() - _t(b).float()).abs().to(t.dtype))

    def absdiff_bwd(self, a, b, g, da, acc_a, db, acc_b):
        s = torch.sign(_t(a).float() - _t(b).float()) * _t(g).float()
        ta, tb = _t(da), _t(db)
        ta.copy_((ta.float() + s if acc_a else s).to(ta.dtype))
        tb.copy_((tb.float() - s if acc_b else -s).to(tb.dtype))

    # -- ViT encoder / FloodViT head passes ------------------------------------------------------
    def layernorm_fwd(self, x, gamma, beta, eps, y, mean=None, rstd=None, copy_out=None):
        xf = x.float()
        mu = xf.mean(1)
        var = ((xf - mu[:, None]) ** 2).mean(1)
        rs = torch.rsqrt(var + eps)
        y.copy_((((xf - mu[:, None]) * rs[:, None]) * gamma + beta).to(y.dtype))
        if mean is not None:
            mean.copy_(mu)
        if rstd is not None:
            rstd.copy_(rs)
        if copy_out is not None:
            copy_out.copy_(x)

    def layernorm_bwd(self, dy, x, mean, rstd, gamma, dx, accumulate_dx, dgamma, dbeta):
        xh = (x.float() - mean[:, None]) * rstd[:, None]
        d = dy.float()
        g = d * gamma
        if dx is not None:
            v = rstd[:, None] * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True))
            dx.copy_((dx.float() + v if accumulate_dx else v).to(dx.dtype))
        if dgamma is not None:
            dgamma += (d * xh).sum(0)
        if dbeta is not None:
            dbeta += d.sum(0)

    @staticmethod
    def _patches(img):
        B, Cc, Hi, Wi = img.shape
        gh, gw = Hi // 16, Wi // 16
        return img.view(B, Cc, gh, 16, gw, 16).permute(0, 2, 4, 3, 5, 1).reshape(B, gh * gw, 256 * Cc)   # b (h w) (p1 p2 c)

    def patchify_ln(self, img, Tp, gamma, beta, eps, out, mean, rstd):
        x = self._patches(img.float())
        B, n, PD = x.shape
        mu = x.mean(2)
        rs = torch.rsqrt(((x - mu[..., None]) ** 2).mean(2) + eps)
        y = (x - mu[..., None]) * rs[..., None] * gamma + beta
        out.view(B, Tp, PD)[:, 1:1 + n] = y.to(out.dtype)
        mean.view(B, Tp)[:, 1:1 + n] = mu
        rstd.view(B, Tp)[:, 1:1 + n] = rs

    def patchify_ln_bwd(self, img, Tp, mean, rstd, dy, dgamma, dbeta):
        x = self._patches(img.float())
        B, n, PD = x.shape
        xh = (x - mean.view(B, Tp)[:, 1:1 + n, None]) * rstd.view(B, Tp)[:, 1:1 + n, None]
        d = dy.view(B, Tp, PD)[:, 1:1 + n].float()
        dgamma += (d * xh).sum((0, 1))
        dbeta += d.sum((0, 1))

    def vit_assemble(self, B, T, Tp, e, cls, pos, x0):
        D = e.shape[1]
        v = torch.zeros(B, Tp, D)
        v[:, 1:T] = e.view(B, Tp, D)[:, 1:T].float()
        v[:, 0] = cls.view(1, D)
        v[:, :T] += pos.view(-1, D)[:T]
        x0.copy_(v.view(B * Tp, D).to(x0.dtype))

    def vit_assemble_bwd(self, B, T, Tp, dx0, de, dcls, dpos):
        D = dx0.shape[1]
        g = dx0.view(B, Tp, D).float()
        dpos.view(-1, D)[:T] = g[:, :T].sum(0)
        dcls.view(-1)[:D] = g[:, 0].sum(0)
        o = torch.zeros(B, Tp, D)
        o[:, 1:T] = g[:, 1:T]
        de.copy_(o.view(B * Tp, D).to(de.dtype))

    def attention_fwd(self, B, T, Tp, heads, dh, qkv, scale, out, probs):
        inner = heads * dh
        q, k, v = [t.view(B, Tp, heads, dh).permute(0, 2, 1, 3)[:, :, :T] for t in qkv.float().view(B, Tp, 3 * inner).split(inner, dim=2)]
        p = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1).to(probs.dtype).float()
        pf = torch.zeros(B, heads, Tp, Tp)
        pf[:, :, :T, :T] = p
        probs.copy_(pf.view(probs.shape).to(probs.dtype))
        o = torch.zeros(B, Tp, heads, dh)
        o[:, :T] = (p @ v).permute(0, 2, 1, 3)
        out.copy_(o.view(B * Tp, inner).to(out.dtype))

    def attention_bwd(self, B, T, Tp, heads, dh, qkv, probs, dout, scale, dqkv, ds_scratch):
        inner = heads * dh
        q, k, v = [t.view(B, Tp, heads, dh).permute(0, 2, 1, 3)[:, :, :T] for t in qkv.float().view(B, Tp, 3 * inner).split(inner, dim=2)]
        p = probs.float().view(B, heads, Tp, Tp)[:, :, :T, :T]
        do = dout.float().view(B, Tp, heads, dh).permute(0, 2, 1, 3)[:, :, :T]
        dv = p.transpose(-1, -2) @ do
        dp = do @ v.transpose(-1, -2)
        ds = (p * (dp - (dp * p).sum(-1, keepdim=True))).to(dqkv.dtype).float()
        dq = ds @ k * scale
        dk = ds.transpose(-1, -2) @ q * scale
        o = torch.zeros(B, Tp, 3, heads, dh)
        for i, t in enumerate((dq, dk, dv)):
            o[:, :T, i] = t.permute(0, 2, 1, 3)
        dqkv.copy_(o.view(B * Tp, 3 * inner).to(dqkv.dtype))

    def gelu_fwd(self, u, h):
        h.copy_(F.gelu(u.float()).to(h.dtype))

    def gelu_bwd(self, u, dh, du):
        x = u.float()
        d = 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * 0.3989422804014327 * torch.exp(-0.5 * x * x)
        du.copy_((dh.float() * d).to(du.dtype))

    def bilinear_up_fwd(self, B, G, Tp, row0, K, Ho, Wo, src, dst):
        Cs = src.shape[1]
        m = src.float().view(B, Tp, Cs)[:, row0:row0 + G * G, :K].reshape(B, G, G, K).permute(0, 3, 1, 2)
        dst.copy_(F.interpolate(m, size=(Ho, Wo), mode="bilinear", align_corners=False))

    def bilinear_up_bwd(self, B, G, Tp, row0, K, Ho, Wo, ddst, dsrc):
        Cs = dsrc.shape[1]
        with torch.enable_grad():      # called from inside autograd.Function.backward, where grad mode is off
            m = torch.zeros(B, K, G, G, requires_grad=True)
            F.interpolate(m, size=(Ho, Wo), mode="bilinear", align_corners=False).backward(ddst.float())
        o = torch.zeros(B, Tp, Cs)
        o[:, row0:row0 + G * G, :K] = m.grad.permute(0, 2, 3, 1).reshape(B, G * G, K)
        dsrc.copy_(o.view(B * Tp, Cs).to(dsrc.dtype))

    def confusion_update(self, pred, labels, K, ignore_index, mat):
        t, p = labels.reshape(-1), pred.reshape(-1).long()
        keep = t != ignore_index
        mat.view(-1).add_(torch.bincount(t[keep] * K + p[keep], minlength=K * K))

    # -- ChangeFormer passes ---------------------------------------------------------------------
    def sar_preprocess(self, raw, out, mean, std, clamp_max):
        """The reference's own ops (dataset/Dataset.py:162-168, :192-198)."""
        img = raw.float()
        if clamp_max:
            img = torch.nan_to_num(torch.clamp(img, min=0.0, max=clamp_max), clamp_max)
        else:
            img = torch.nan_to_num(img, 200)
        out.copy_((img - mean.view(1, -1, 1, 1)) / std.view(1, -1, 1, 1))

    def confusion_update_grouped(self, pred, labels, K, ignore_index, mat, key_a=None, mat_a=None, key_b=None, mat_b=None):
        for s in range(labels.shape[0]):
            one = torch.zeros(K, K, dtype=torch.int64)
            self.confusion_update(pred[s].reshape(-1), labels[s].reshape(-1), K, ignore_index, one)
            if mat is not None:
                mat += one
            for key, m in ((key_a, mat_a), (key_b, mat_b)):
                if m is not None and 0 <= int(key[s]) < m.shape[0]:
                    m[int(key[s])] += one

    @staticmethod
    def _w_oihw(weight, k, cout, cin):
        return weight.float().view(k * k, cout, cin).permute(1, 2, 0).reshape(cout, cin, k, k)

    def conv2d_strided(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, src, weight, bias, dst):
        x = _t(src).float().permute(0, 3, 1, 2)
        y = F.conv2d(x, self._w_oihw(weight, ksize, dst.C, src.C), None if bias is None else bias.float(), stride=stride, padding=pad)
        t = _t(dst)
        t.copy_(y.permute(0, 2, 3, 1).to(t.dtype))

    def conv2d_strided_dgrad(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, dy, weight, dx, accumulate=False):
        g = _t(dy).float().permute(0, 3, 1, 2)
        v = torch.nn.grad.conv2d_input((N, dx.C, Hi, Wi), self._w_oihw(weight, ksize, dy.C, dx.C), g, stride=stride, padding=pad).permute(0, 2, 3, 1)
        t = _t(dx)
        t.copy_((t.float() + v if accumulate else v).to(t.dtype))

    def conv2d_strided_wgrad(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, x, dy, dw, accumulate=False):
        xx, g = _t(x).float().permute(0, 3, 1, 2), _t(dy).float().permute(0, 3, 1, 2)
        gw = torch.nn.grad.conv2d_weight(xx, (dy.C, x.C, ksize, ksize), g, stride=stride, padding=pad)
        gw = gw.reshape(dy.C, x.C, ksize * ksize).permute(2, 0, 1).reshape(-1)
        out = dw.reshape(-1)[: gw.numel()]
        out.copy_(gw + out if accumulate else gw)

    def im2col(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, x, col, Kp):
        xx = F.pad(_t(x).float(), (0, 0, pad, pad, pad, pad))                       # [N, Hi+2p, Wi+2p, C]
        c = col.view(N, Ho, Wo, Kp)
        Cn = x.C
        for u in range(ksize):
            for v in range(ksize):
                t = u * ksize + v
                c[..., t * Cn:(t + 1) * Cn] = xx[:, u:u + (Ho - 1) * stride + 1:stride, v:v + (Wo - 1) * stride + 1:stride, :].to(col.dtype)

    def col2im(self, N, Hi, Wi, Ho, Wo, ksize, stride, pad, dcol, Kp, dx, accumulate=False):
        Cn = dx.C
        acc = torch.zeros(N, Hi + 2 * pad + ksize, Wi + 2 * pad + ksize, Cn)
        c = dcol.view(N, Ho, Wo, Kp).float()
        for u in range(ksize):
            for v in range(ksize):
                t = u * ksize + v
                acc[:, u:u + (Ho - 1) * stride + 1:stride, v:v + (Wo - 1) * stride + 1:stride, :] += c[..., t * Cn:(t + 1) * Cn]
        g = acc[:, pad:pad + Hi, pad:pad + Wi, :]
        t_ = _t(dx)
        t_.copy_((t_.float() + g if accumulate else g).to(t_.dtype))

    @staticmethod
    def _xa_split(B, Nq, Nk, heads, dh, q, kv):
        inner = heads * dh
        qq = q.float()[:, :inner].reshape(B, Nq, heads, dh).permute(0, 2, 1, 3)
        k = kv.float()[:, :inner].reshape(B, Nk, heads, dh).permute(0, 2, 1, 3)
        v = kv.float()[:, inner:2 * inner].reshape(B, Nk, heads, dh).permute(0, 2, 1, 3)
        return qq, k, v

    # stateless RNG of the stochastic regularisers: the numpy restatement lives in oracle/changeformer_oracle.py
    def keep_factors(self, n, p, seed, step, site):
        from oracle.changeformer_oracle import keep_factors
        return keep_factors(n, p, seed, 0 if step is None else int(step.item()), site)

    def dropout_apply(self, x, y, p, seed, step, site):
        f = self.keep_factors(x.numel(), p, seed, step, site).view(x.shape)
        y.copy_((x.float() * f).to(y.dtype))

    def _branch_f(self, t, per_sample, p, droppath, seed, step, site):
        f = self.keep_factors(t.numel(), p, seed, step, site) if p > 0 else torch.ones(t.numel())
        if droppath is not None:
            f = f * droppath.reshape(-1)[torch.arange(t.numel()) // per_sample]
        return f.view(t.shape)

    def branch_add(self, x, t, per_sample, p, droppath, seed, step, site):
        x.copy_((x.float() + self._branch_f(t, per_sample, p, droppath, seed, step, site) * t.float()).to(x.dtype))

    def branch_scale(self, dx, dt, per_sample, p, droppath, seed, step, site):
        dt.copy_((self._branch_f(dx, per_sample, p, droppath, seed, step, site) * dx.float()).to(dt.dtype))

    def xattention_fwd(self, B, Nq, Nk, heads, dh, q, kv, scale, out, probs, pdrop=0.0, seed=0, step=None, site=0):
        qq, k, v = self._xa_split(B, Nq, Nk, heads, dh, q, kv)
        p = torch.softmax(qq @ k.transpose(-1, -2) * scale, dim=-1).to(probs.dtype)
        probs.copy_(p.reshape(probs.shape))
        pd = p.float()
        if pdrop > 0:
            pd = pd * self.keep_factors(pd.numel(), pdrop, seed, step, site).view(pd.shape)
        out[:, :heads * dh] = (pd @ v).permute(0, 2, 1, 3).reshape(B * Nq, heads * dh).to(out.dtype)

    def xattention_bwd(self, B, Nq, Nk, heads, dh, q, kv, probs, dout, scale, dq, dkv_f32, pdrop=0.0, seed=0, step=None, site=0):
        inner = heads * dh
        qq, k, v = self._xa_split(B, Nq, Nk, heads, dh, q, kv)
        p = probs.float().view(B, heads, Nq, Nk)
        do = dout.float()[:, :inner].reshape(B, Nq, heads, dh).permute(0, 2, 1, 3)
        kf = self.keep_factors(p.numel(), pdrop, seed, step, site).view(p.shape) if pdrop > 0 else torch.ones_like(p)
        dv = (p * kf).transpose(-1, -2) @ do
        dp = (do @ v.transpose(-1, -2)) * kf
        ds = p * (dp - (dp * p).sum(-1, keepdim=True)) * scale
        dq[:, :inner] = (ds @ k).permute(0, 2, 1, 3).reshape(B * Nq, inner).to(dq.dtype)
        dk = ds.transpose(-1, -2) @ qq
        dkv_f32.view(B * Nk, 2 * inner)[:, :inner] = dk.permute(0, 2, 1, 3).reshape(B * Nk, inner)
        dkv_f32.view(B * Nk, 2 * inner)[:, inner:] = dv.permute(0, 2, 1, 3).reshape(B * Nk, inner)

    def dwconv3x3_fwd(self, N, H, W, x, w9, bias, y):
        Cn = x.shape[-1]
        xx = x.float().view(N, H, W, Cn).permute(0, 3, 1, 2)
        w = w9.float().view(9, Cn).t().reshape(Cn, 1, 3, 3)
        o = F.conv2d(xx, w, None if bias is None else bias.float(), padding=1, groups=Cn).permute(0, 2, 3, 1)
        y.copy_(o.reshape(y.shape).to(y.dtype))

    def dwconv3x3_bwd(self, N, H, W, x, dy, w9, dx, dw9, dbias):
        Cn = x.shape[-1]
        with torch.enable_grad():
            xx = x.float().view(N, H, W, Cn).permute(0, 3, 1, 2).detach().requires_grad_(True)
            w = w9.float().view(9, Cn).t().reshape(Cn, 1, 3, 3).detach().requires_grad_(True)
            o = F.conv2d(xx, w, None, padding=1, groups=Cn)
            gx, gw = torch.autograd.grad(o, (xx, w), dy.float().view(N, H, W, Cn).permute(0, 3, 1, 2))
        dx.copy_(gx.permute(0, 2, 3, 1).reshape(dx.shape).to(dx.dtype))
        dw9.view(9, Cn).add_(gw.reshape(Cn, 9).t())
        if dbias is not None:
            dbias.add_(dy.float().reshape(-1, Cn).sum(0))

    def bilinear_nhwc_fwd(self, N, Hi, Wi, Ho, Wo, src, dst, accumulate=False):
        Cn = src.shape[-1]
        v = F.interpolate(src.float().view(N, Hi, Wi, Cn).permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
        v = v.reshape(dst.shape)
        dst.copy_((dst.float() + v if accumulate else v).to(dst.dtype))

    def bilinear_nhwc_bwd(self, N, Hi, Wi, Ho, Wo, ddst, dsrc, accumulate=False):
        Cn = ddst.shape[-1]
        with torch.enable_grad():
            m = torch.zeros(N, Cn, Hi, Wi, requires_grad=True)
            F.interpolate(m, size=(Ho, Wo), mode="bilinear", align_corners=False).backward(ddst.float().view(N, Ho, Wo, Cn).permute(0, 3, 1, 2))
        v = m.grad.permute(0, 2, 3, 1).reshape(dsrc.shape)
        dsrc.copy_((dsrc.float() + v if accumulate else v).to(dsrc.dtype))

    def relu_fwd(self, x, y):
        y.copy_(F.relu(x.float()).to(y.dtype))

    def relu_bwd(self, r, g, dx):
        dx.copy_((g.float() * (r.float() > 0)).to(dx.dtype))

    def sigmoid_head_fwd(self, z, K, out):
        out.copy_(torch.sigmoid(_t(z).float()[..., :K].permute(0, 3, 1, 2)))

    def sigmoid_head_bwd(self, out, dout, K, dz):
        t = _t(dz)
        t.zero_()
        t[..., :K].copy_((dout * out * (1 - out)).permute(0, 2, 3, 1).to(t.dtype))

    def sgd_step(self, p, g, buf, lr, momentum, wd, grad_scale):
        gg = g * grad_scale + wd * p
        buf.mul_(momentum).add_(gg)
        p.add_(buf, alpha=-lr)

    def adaptive_avgpool_fwd(self, src, S, dst):
        x = _t(src).float().permute(0, 3, 1, 2)
        dst.copy_(F.adaptive_avg_pool2d(x, S).permute(0, 2, 3, 1).reshape(dst.shape).to(dst.dtype))

    def adaptive_avgpool_bwd(self, ddst, S, dsrc, accumulate=False):
        t = _t(dsrc)
        with torch.enable_grad():
            m = torch.zeros(dsrc.N, dsrc.C, dsrc.H, dsrc.W, requires_grad=True)
            F.adaptive_avg_pool2d(m, S).backward(ddst.float().view(dsrc.N, S, S, dsrc.C).permute(0, 3, 1, 2))
        v = m.grad.permute(0, 2, 3, 1)
        t.copy_((t.float() + v if accumulate else v).to(t.dtype))
