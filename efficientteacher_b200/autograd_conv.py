"""torch.autograd bridge for the tcgen05 convolutions: forward = implicit-GEMM conv, backward = dgrad (same kernel,
transposed taps) + wgrad (pixel-reduction GEMM).  Tensors stay NHWC bf16 in HBM and are exposed to torch as
channels_last NCHW views (zero copy), so torch's BatchNorm/SiLU/cat/upsample/maxpool can sit between the convs while
the dense contractions (97% of the step's FLOPs) run on the hand-written kernels.

Replaces cuDNN fprop/dgrad/wgrad behind Conv.forward (reference models/backbone/common.py:480-481), the Detect 1x1
convs (models/head/yolov5_head.py:55) and netD.conv1 (models/detector/yolo_ssod.py:228).
"""
import torch

from . import _lib
from . import convops as co


ACCUMULATE_INTO_GRAD = True


def _as_nhwc(t, C_):
    """(view [N,H,W,C] bf16 whose channel stride may exceed C, channel_stride) for an NCHW-shaped tensor; copies only if
    the layout is not NHWC.  A channel-slice of a channels_last tensor (what torch.cat's backward hands out) is used in
    place: the kernels take the pixel stride separately and never touch channels outside the slice."""
    N, C2, H, W = t.shape
    assert C2 == C_
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    sN, sC, sH, sW = t.stride()
    cs = sW
    ok = sC == 1 and cs >= C_ and cs % 8 == 0 and sH == W * cs and sN == H * W * cs and t.data_ptr() % 16 == 0
    if not ok:
        t = t.contiguous(memory_format=torch.channels_last)
        cs = C_
        if t.stride() != (H * W * C_, 1, W * C_, C_):     # degenerate shapes (H=W=1 ...): force the NHWC strides
            t = t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    return t.permute(0, 2, 3, 1), cs


def _empty_cl(N, C_, H, W, device):
    """NCHW-shaped bf16 tensor with channels_last strides (physically NHWC); returned from the Functions as a real
    tensor (not a view) so downstream in-place ops (ReLU(inplace)) are legal."""
    return torch.empty((N, C_, H, W), dtype=torch.bfloat16, device=device, memory_format=torch.channels_last)


def _nhwc_of(t):
    N, C_, H, W = t.shape
    v = t.permute(0, 2, 3, 1)
    if v.stride() != (H * W * C_, W * C_, C_, 1):       # degenerate sizes: channels_last strides are ambiguous
        raise RuntimeError("unexpected channels_last strides %s for %s" % (t.stride(), tuple(t.shape)))
    return v


class ConvFn(torch.autograd.Function):
    """y = conv2d(x, w) (no bias, no activation), bf16 channels_last in / out."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        Cout, Cin, k, _ = weight.shape
        xb, xcs = _as_nhwc(x, Cin)
        wp = co.pack_weight(weight)
        N, _, H, W = x.shape
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = _empty_cl(N, Cout, Ho, Wo, x.device)
        co.conv_fwd(xb, wp, Cin, Cout, k, stride, pad, None, None, None, x_cstride=xcs, out=_nhwc_of(y))
        ctx.save_for_backward(x, weight)
        ctx.geom = (stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        stride, pad = ctx.geom
        Cout, Cin, k, _ = weight.shape
        N, _, H, W = x.shape
        dyb, dycs = _as_nhwc(dy, Cout)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wd = co.pack_weight_dgrad(weight, stride, pad)
            dx = _empty_cl(N, Cin, H, W, dy.device)
            co.conv_dgrad(dyb, wd, N, H, W, Cin, Cout, k, stride, pad, dy_cstride=dycs, out=_nhwc_of(dx))
        if ctx.needs_input_grad[1]:
            xb, xcs = _as_nhwc(x, Cin)
            dw = co.conv_wgrad(xb, dyb, Cin, Cout, k, stride, pad, x_cstride=xcs, dy_cstride=dycs)
        return dx, dw, None, None


class ConvBnActFn(torch.autograd.Function):
    """a = act(BatchNorm_train(conv2d(x, w))) -- the whole Conv module (common.py:480-481) in training mode:
    tcgen05 conv -> per-channel batch statistics -> fused normalise+SiLU; backward = fused SiLU'/BN backward (2 passes)
    -> dgrad + wgrad.  Saves x, the raw conv output and [4,C] statistics (not the normalised tensor)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, stride, pad, eps, momentum, act, is_stem,
                wp=None, wd=None):
        Cout = weight.shape[0]
        ctx.wd = wd
        if is_stem:
            xb = co.stem_im2col(x.float(), 1.0)               # [N,H/2,W/2,128]; saved instead of the image
            xcs, Cin, k, st, pd = 128, 128, 1, 1, 0
            if wp is None:
                wp = co.pack_stem_weight(weight)
            N, Ho, Wo = xb.shape[0], xb.shape[1], xb.shape[2]
        else:
            Cin, k = weight.shape[1], weight.shape[2]
            xb, xcs = _as_nhwc(x, Cin)
            st, pd = stride, pad
            if wp is None:
                wp = co.pack_weight(weight)
            N, _, H, W = x.shape
            Ho, Wo = (H + 2 * pd - k) // st + 1, (W + 2 * pd - k) // st + 1
        y = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
        co.conv_fwd(xb, wp, Cin, Cout, k, st, pd, None, None, None, x_cstride=xcs, out=y)
        a = _empty_cl(N, Cout, Ho, Wo, x.device)
        _, stats = co.bn_forward(y, Cout, gamma.detach(), beta.detach(), running_mean, running_var, eps, momentum, act, out=_nhwc_of(a))
        ctx.save_for_backward(xb if is_stem else x, weight, y, stats)
        ctx.meta = (stride, pad, act, is_stem)
        return a

    @staticmethod
    def backward(ctx, da):
        xs, weight, y, stats = ctx.saved_tensors
        stride, pad, act, is_stem = ctx.meta
        Cout = weight.shape[0]
        dab, dacs = _as_nhwc(da, Cout)
        dy, dgamma, dbeta = co.bn_backward(dab, y, Cout, stats, act, da_cstride=dacs)
        dx = dw = None
        # gradient arena: when p.grad already exists (trainer.GradArena) the wgrad is added into it in place and autograd
        # gets None for the weight (no separate AccumulateGrad add pass, no temporary)
        tgt = weight.grad if (ACCUMULATE_INTO_GRAD and weight.grad is not None and weight.grad.is_contiguous()) else None
        if is_stem:
            dw = co.conv_wgrad(xs, dy, 128, Cout, 1, 1, 0, stem=True, accumulate_into=tgt)
        else:
            Cin, k = weight.shape[1], weight.shape[2]
            N, _, H, W = xs.shape
            if ctx.needs_input_grad[0]:
                dx = _empty_cl(N, Cin, H, W, da.device)
                wd = ctx.wd if ctx.wd is not None else co.pack_weight_dgrad(weight, stride, pad)
                co.conv_dgrad(dy, wd, N, H, W, Cin, Cout, k, stride, pad, out=_nhwc_of(dx))
            xb, xcs = _as_nhwc(xs, Cin)
            dw = co.conv_wgrad(xb, dy, Cin, Cout, k, stride, pad, x_cstride=xcs, accumulate_into=tgt)
        if tgt is not None:
            dw = None
        return dx, dw, dgamma, dbeta, None, None, None, None, None, None, None, None, None, None


class StemFn(torch.autograd.Function):
    """The 6x6 s2 p2 stem on the raw fp32 NCHW image: im2col (K=108 padded to 128) + pointwise GEMM.  No input grad."""

    @staticmethod
    def forward(ctx, x, weight):
        col = co.stem_im2col(x, 1.0)
        N, H2, W2, _ = col.shape
        y = _empty_cl(N, weight.shape[0], H2, W2, x.device)
        co.conv_fwd(col, co.pack_stem_weight(weight), 128, weight.shape[0], 1, 1, 0, None, None, None, out=_nhwc_of(y))
        ctx.save_for_backward(col)
        ctx.cout = weight.shape[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        dyb, dycs = _as_nhwc(dy, ctx.cout)
        dw = co.conv_wgrad(col, dyb, 128, ctx.cout, 1, 1, 0, dy_cstride=dycs, stem=True)
        return None, dw


class DetectConvFn(torch.autograd.Function):
    """Detect's 1x1 conv + bias, emitting fp32 logits directly in the train layout [N,na,ny,nx,no]
    (the view/permute/contiguous of models/head/yolov5_head.py:66 is fused into the epilogue)."""

    @staticmethod
    def forward(ctx, x, weight, bias, na, no):
        Cout, Cin = weight.shape[0], weight.shape[1]
        xb, xcs = _as_nhwc(x, Cin)
        N, H, W, _ = xb.shape
        out = torch.empty((N, na, H, W, no), dtype=torch.float32, device=x.device)
        co.conv_fwd(xb, co.pack_weight(weight), Cin, Cout, 1, 1, 0, None, bias.detach().float().contiguous(), None, x_cstride=xcs,
                    det_out=out, det_no=no)
        ctx.save_for_backward(x, weight)
        ctx.meta = (na, no)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        na, no = ctx.meta
        Cout, Cin = weight.shape[0], weight.shape[1]
        N, _, H, W, _ = g.shape
        # [N,na,H,W,no] fp32 -> NHWC bf16 [N,H,W,256] (channel c = a*no+o; padded to a multiple of 8 channels)
        cpad = (Cout + 7) // 8 * 8
        dyb = torch.zeros((N, H, W, cpad), dtype=torch.bfloat16, device=g.device)
        dyb[..., :Cout] = g.permute(0, 2, 3, 1, 4).reshape(N, H, W, Cout)
        db = g.sum((0, 2, 3)).reshape(-1) if ctx.needs_input_grad[2] else None
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dgrad needs K = Cout to be a multiple of 64: use the zero-padded 256-channel view of dy and of W^T
            kpad = (Cout + 63) // 64 * 64
            if kpad != cpad:
                d2 = torch.zeros((N, H, W, kpad), dtype=torch.bfloat16, device=g.device)
                d2[..., :Cout] = dyb[..., :Cout]
            else:
                d2 = dyb
            wpad = torch.zeros((kpad, Cin, 1, 1), dtype=torch.float32, device=g.device)
            wpad[:Cout] = weight.detach().float()
            dx = _empty_cl(N, Cin, H, W, g.device)
            co.conv_dgrad(d2, co.pack_weight_dgrad(wpad, 1, 0), N, H, W, Cin, kpad, 1, 1, 0, out=_nhwc_of(dx))
        if ctx.needs_input_grad[1]:
            xb, xcs = _as_nhwc(x, Cin)
            dw = co.conv_wgrad(xb, dyb, Cin, Cout, 1, 1, 0, x_cstride=xcs)
        return dx, dw, db, None, None


def conv2d_native(x, weight, stride, pad):
    _lib.require_cuda(x, weight)
    return ConvFn.apply(x, weight, stride, pad)
