"""DomainLoss / TargetLoss with the reference's surface (models/loss/loss.py:375-421; DomainFocalLoss :312-372 with its
defaults class_num=2, gamma=2, alpha=1, softmax branch): 0.5 * mean over all positions of the three netD maps of
-(1 - p)^2 * log p, p = softmax(logits)[domain label] (0 = labeled / source, 1 = unlabeled / target).

Forward and backward are one fused kernel each over all three levels (csrc/tail.cu: etb_domain_focal_fwd / _bwd) instead of
permute/reshape/cat/softmax/scatter/pow/log/mean and their autograd twins.  Unlike the reference (loss.py:392,418 hard-code
`.cuda()`) the labels never exist as tensors."""
import ctypes as C

import torch

from . import _lib
from ._lib import EtbFocalParams


def _maps_as_m2(feature):
    """each [B,2,H,W] map as a contiguous fp32 [B*H*W, 2] run (zero copy for the native netD output: an NCHW-shaped view of a
    [B,H,W,2] buffer, or a batch slice of one)"""
    out = []
    for f in feature:
        if f.dim() != 4 or f.shape[1] != 2:
            raise RuntimeError("domain loss expects netD maps [B,2,H,W], got %s" % (tuple(f.shape),))
        v = f.permute(0, 2, 3, 1)
        if v.dtype != torch.float32 or not v.is_contiguous():
            v = v.float().contiguous()
        out.append(v)
    return out


class _DomainFocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, label, *feature):
        _lib.require_cuda(*feature)
        maps = _maps_as_m2(feature)
        lib = _lib.lib()
        dev = maps[0].device
        fp = EtbFocalParams()
        fp.nl, fp.label = len(maps), int(label)
        for l, m in enumerate(maps):
            fp.x[l], fp.M[l] = m.data_ptr(), m.numel() // 2
        ws = torch.empty(int(lib.etb_domain_focal_workspace_bytes()), dtype=torch.uint8, device=dev)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.check(lib.etb_domain_focal_fwd(C.byref(fp), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "etb_domain_focal_fwd")
        ctx.maps, ctx.feature, ctx.label = maps, feature, int(label)
        return out

    @staticmethod
    def backward(ctx, g):
        from .autograd_conv import grad_buffer_for
        maps, feature = ctx.maps, ctx.feature
        fp = EtbFocalParams()
        fp.nl, fp.label = len(maps), ctx.label
        grads = []
        for l, (m, f) in enumerate(zip(maps, feature)):
            gf = grad_buffer_for(f)                       # [B,2,H,W]-shaped; for the native layout a contiguous [M,2] run
            gv = gf.permute(0, 2, 3, 1)
            direct = gv.dtype == torch.float32 and gv.is_contiguous()
            dst = gv if direct else torch.empty_like(m)
            fp.x[l], fp.dx[l], fp.M[l] = m.data_ptr(), dst.data_ptr(), m.numel() // 2
            grads.append((gf, gv, dst, direct))
        _lib.check(_lib.lib().etb_domain_focal_bwd(C.byref(fp), _lib.ptr(g.float().contiguous()), _lib.stream_ptr()), "etb_domain_focal_bwd")
        outs = []
        for gf, gv, dst, direct in grads:
            if not direct:
                gv.copy_(dst)
            outs.append(gf)
        return (None, *outs)


def domain_focal_loss(feature, label):
    """0.5 * DomainFocalLoss(class_num=2)(cat of the three maps, label) -- reference loss.py:385-393 / :411-420"""
    return _DomainFocalFn.apply(int(label), *feature)[0]


class DomainLoss:        # reference models/loss/loss.py:398-421 (source / labeled batch: domain label 0)
    def __call__(self, feature):
        return domain_focal_loss(feature, 0)


class TargetLoss:        # reference models/loss/loss.py:375-395 (target / unlabeled batch: domain label 1)
    def __call__(self, feature):
        return domain_focal_loss(feature, 1)
