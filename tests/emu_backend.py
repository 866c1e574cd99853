"""TEST INFRASTRUCTURE ONLY: a plain-PyTorch emulation of the C-ABI entry points (include/b200em.h), one method
per entry point with the contract stated in the header.  It lets the CPU test-suite check the HOST logic of the
product -- the kernel schedule in torch-em_b200/engine.py: which buffer feeds which kernel, forward and backward --
against the oracle / golden vectors without a GPU.  The product never imports this file; its only backend is
torch-em_b200/backend.py (CUDA), which raises on CPU tensors.
"""
import torch
import torch.nn.functional as F


class _Pack:
    pass


def _bcast(v):  # (N, C) -> (N, 1, 1, 1, C)
    return v[:, None, None, None, :]


def _ncdhw(t):
    return t.permute(0, 4, 1, 2, 3)


class TorchEmuBackend:
    name = "torch-emulation"

    def pack(self, key, w, lazy_dgrad=False):
        p = _Pack()
        p.w = w.detach()
        return p

    def to_ndhwc(self, x, y):
        y.copy_(x.permute(0, 2, 3, 4, 1))

    def channel_sums(self, x, sums):
        xf = x.float()
        sums[:, :, 0] += xf.sum((1, 2, 3))
        sums[:, :, 1] += (xf * xf).sum((1, 2, 3))

    def channel_dot_sums(self, g, x, sums):
        gf, xf = g.float(), x.float()
        sums[:, :, 0] += gf.sum((1, 2, 3))
        sums[:, :, 1] += (gf * xf).sum((1, 2, 3))

    def norm_finalize(self, sums, S, groups, gamma, beta, eps):
        N, C, _ = sums.shape
        cpg = C // groups
        s = sums.double().reshape(N, groups, cpg, 2).sum(2)
        cnt = float(S * cpg)
        mean = s[..., 0] / cnt
        var = (s[..., 1] / cnt - mean * mean).clamp(min=0)
        rstd = 1.0 / torch.sqrt(var + eps)
        mean = mean.repeat_interleave(cpg, 1).float()
        rstd = rstd.repeat_interleave(cpg, 1).float()
        ga = gamma.detach() if gamma is not None else torch.ones(C)
        be = beta.detach() if beta is not None else torch.zeros(C)
        ss = torch.stack([rstd * ga, be - mean * rstd * ga], -1)
        mr = torch.stack([mean, rstd], -1)
        return ss, mr

    def norm_bwd_finalize(self, dsums, mr, gamma, S, groups, dgamma, dbeta):
        N, C, _ = dsums.shape
        cpg = C // groups
        mean, rstd = mr[..., 0], mr[..., 1]
        ga = gamma.detach() if gamma is not None else torch.ones(C)
        sg, sgx = dsums[..., 0], dsums[..., 1]
        sgxh = rstd * (sgx - mean * sg)
        if dgamma is not None:
            dgamma += sgxh.sum(0)
        if dbeta is not None:
            dbeta += sg.sum(0)
        cnt = float(S * cpg)
        m1 = ((ga * sg).reshape(N, groups, cpg).sum(2) / cnt).repeat_interleave(cpg, 1)
        m2 = ((ga * sgxh).reshape(N, groups, cpg).sum(2) / cnt).repeat_interleave(cpg, 1)
        return torch.stack([rstd * ga, -rstd * rstd * m2, rstd * (mean * rstd * m2 - m1)], -1)

    def norm_bwd_apply(self, g, x, coef, add, out, relu_mask):
        t = g.float()
        if coef is not None:
            t = _bcast(coef[..., 0]) * g.float() + _bcast(coef[..., 1]) * x.float() + _bcast(coef[..., 2])
        if add is not None:
            t = t + add.float()
        if relu_mask:
            t = t * (x.float() > 0)
        out.copy_(t)

    def _xhat(self, x, in_ss):
        xf = x.float()
        if in_ss is not None:
            xf = xf * _bcast(in_ss[..., 0]) + _bcast(in_ss[..., 1])
        return _ncdhw(xf)

    def conv(self, x, in_ss, pack, bias, y, sums, kernel, relu, dgrad, dot_x=None):
        pad = tuple(k // 2 for k in kernel)
        xh = self._xhat(x, in_ss)
        if dgrad:
            r = F.conv_transpose3d(xh, pack.w, None, padding=pad)
        else:
            r = F.conv3d(xh, pack.w, bias.detach() if bias is not None else None, padding=pad)
        if relu:
            r = F.relu(r)
        y.copy_(r.permute(0, 2, 3, 4, 1))
        if sums is not None:
            if dot_x is not None:
                self.channel_dot_sums(y, dot_x, sums)
            else:
                self.channel_sums(y, sums)
        return None

    def wgrad(self, x, in_ss, dz, dw, db, kernel, aux=None):
        pad = tuple(k // 2 for k in kernel)
        xh = self._xhat(x, in_ss).contiguous()
        dw += torch.nn.grad.conv3d_weight(xh, dw.shape, _ncdhw(dz.float()).contiguous(), padding=pad)
        if db is not None:
            db += dz.float().sum((0, 1, 2, 3))

    def maxpool_fwd(self, x, y, f, sums):
        r = F.max_pool3d(_ncdhw(x.float()), kernel_size=list(f), stride=list(f))
        y.copy_(r.permute(0, 2, 3, 4, 1))
        if sums is not None:
            self.channel_sums(y, sums)

    def maxpool_bwd(self, x, dp, add, out, f, relu_mask, coef=None):
        if coef is not None:
            add = _bcast(coef[..., 0]) * add.float() + _bcast(coef[..., 1]) * x.float() + _bcast(coef[..., 2])
        with torch.enable_grad():
            xf = _ncdhw(x.float()).detach().clone().requires_grad_(True)
            r = F.max_pool3d(xf, kernel_size=list(f), stride=list(f))
            (gx,) = torch.autograd.grad(r, xf, _ncdhw(dp.float()))
        t = gx.permute(0, 2, 3, 4, 1)
        if add is not None:
            t = t + add.float()
        if relu_mask:
            t = t * (x.float() > 0)
        out.copy_(t)

    def upsample_fwd(self, x, y, f, sums):
        r = F.interpolate(_ncdhw(x.float()), scale_factor=[float(s) for s in f], mode="trilinear", align_corners=False)
        y.copy_(r.permute(0, 2, 3, 4, 1))
        if sums is not None:
            self.channel_sums(y, sums)

    def fused_up_bwd_ok(self, dy, f):
        return True

    def upsample_bwd(self, dy, dx, f, zlow=None, coef=None):
        if coef is not None:
            up = F.interpolate(_ncdhw(zlow.float()), scale_factor=[float(s) for s in f], mode="trilinear", align_corners=False)
            dy = _bcast(coef[..., 0]) * dy.float() + _bcast(coef[..., 1]) * up.permute(0, 2, 3, 4, 1) + _bcast(coef[..., 2])
        with torch.enable_grad():
            z = torch.zeros(_ncdhw(dx).shape, dtype=torch.float32, requires_grad=True)
            r = F.interpolate(z, scale_factor=[float(s) for s in f], mode="trilinear", align_corners=False)
            (g,) = torch.autograd.grad(r, z, _ncdhw(dy.float()))
        dx.copy_(g.permute(0, 2, 3, 4, 1))

    @staticmethod
    def _act(v, act):
        if act == "Sigmoid":
            return torch.sigmoid(v)
        if act == "ReLU":
            return F.relu(v)
        if act == "Tanh":
            return torch.tanh(v)
        return v

    def head_fwd(self, x, w, b, out, act):
        r = F.conv3d(_ncdhw(x.float()), w.detach(), b.detach())
        out.copy_(self._act(r, act))

    def head_bwd(self, grad_out, out, x, w, dx, dw, db, act, relu_mask):
        if act == "Sigmoid":
            d = out * (1 - out)
        elif act == "ReLU":
            d = (out > 0).float()
        elif act == "Tanh":
            d = 1 - out * out
        else:
            d = torch.ones_like(out)
        dz = grad_out * d                                   # (N, Cout, D, H, W)
        xf = _ncdhw(x.float())
        wm = w.detach().reshape(w.shape[0], w.shape[1])
        dw += torch.einsum("nodhw,ncdhw->oc", dz, xf).reshape(dw.shape)
        db += dz.sum((0, 2, 3, 4))
        if dx is not None:
            t = torch.einsum("nodhw,oc->ndhwc", dz, wm)
            if relu_mask:
                t = t * (x.float() > 0)
            dx.copy_(t)
