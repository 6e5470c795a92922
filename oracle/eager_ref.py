"""TEST INFRASTRUCTURE / BASELINE ONLY -- the SSOD step the way the reference executes it on a GPU with stock PyTorch:
eager ATen / cuDNN / torchvision ops under autocast, a Python loop per image around torchvision.ops.nms, the pseudo-label
transform on the host in numpy, ~40 tiny tensor ops per level in build_targets, ~150 launches per loss, torch.optim.SGD and
per-tensor EMA updates.  It is what SURVEY.md 2.1 / 8(d) call "the GPU-eager reference": the kernel-level bar the native
step has to beat on the same box.  bench.py times it as `gpu_eager_baseline`; tests/test_oracle_golden.py checks its
pieces (on the CPU, where torch ops are deterministic) against oracle/port.py and the golden vectors.

Restated from (paths relative to the reference root), nothing copied:
  trainer/ssod_trainer.py:587-680,458-488 (step)        utils/general.py:887-992 (non_max_suppression_ssod)
  models/assigner/yolo_anchor_assigner.py:319-372,640-697  models/loss/loss.py:138-208, models/loss/ssod/ssod_loss.py:130-288
  utils/metrics.py:207-249 (CIoU)                        utils/torch_utils.py:328-338,405-416 (EMA loops)
Where the reference's own code is slower than a vectorised restatement (output_to_target_ssod's per-detection Python loop,
select_targets' one `.cpu()` per row) the faster form is used: the baseline errs on the fast side.
The product (efficientteacher_b200) never imports this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import port
from .trunk_ref import TrunkRef

STRIDES = (8, 16, 32)


def build_targets_t(targets, anchors_grid, shapes, anchor_t=4.0, with_score=False):
    """yolo_anchor_assigner.py:319-372 (with_score: :640-697) as the reference runs it: tensor ops on targets' device.
    targets [nt, 6|7] (img, cls, x, y, w, h[, score]) -> per level (b, a, gj, gi, tbox[n,4], anch[n,2], tcls[n], tscore[n])."""
    dev = targets.device
    na, nt = 3, targets.shape[0]
    ncol = 8 if with_score else 7
    ai = torch.arange(na, device=dev, dtype=torch.float32).view(na, 1).repeat(1, nt)
    t_all = torch.cat((targets[:, :ncol - 1].float().repeat(na, 1, 1), ai[:, :, None]), 2)        # :333-334
    g = 0.5
    off = torch.tensor([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]], device=dev, dtype=torch.float32) * g
    out = []
    for l, (ny, nx) in enumerate(shapes):
        anchors = torch.as_tensor(anchors_grid[l], dtype=torch.float32, device=dev).view(na, 2)
        gain = torch.ones(ncol, device=dev)
        gain[2:6] = torch.tensor([nx, ny, nx, ny], device=dev, dtype=torch.float32)
        t = t_all * gain
        if nt:
            r = t[:, :, 4:6] / anchors[:, None]
            j = torch.max(r, 1.0 / r).max(2)[0] < anchor_t                                        # :344-346
            t = t[j]
            gxy = t[:, 2:4]
            gxi = gain[[2, 3]] - gxy
            jk = ((gxy % 1.0 < g) & (gxy > 1.0)).T
            lm = ((gxi % 1.0 < g) & (gxi > 1.0)).T
            sel = torch.stack((torch.ones_like(jk[0]), jk[0], jk[1], lm[0], lm[1]))
            t = t.repeat((5, 1, 1))[sel]
            offsets = (torch.zeros_like(gxy)[None] + off[:, None])[sel]
        else:
            t = t_all[0]
            offsets = 0
        b, c = t[:, :2].long().T
        gxy, gwh = t[:, 2:4], t[:, 4:6]
        gij = (gxy - offsets).long()
        gi, gj = gij.T
        a = t[:, -1].long()
        tscore = t[:, 6] if with_score else torch.zeros_like(t[:, 0])
        out.append((b, a, gj.clamp(0, ny - 1), gi.clamp(0, nx - 1), torch.cat((gxy - gij, gwh), 1), anchors[a], c, tscore))
    return out


def ciou_t(b1, b2, eps=1e-7):
    """utils/metrics.py:207-249, CIoU branch, xywh; b1 [n,4], b2 [n,4]."""
    return port.ciou(b1, b2, eps)          # pure torch ops: runs on whatever device the tensors live on


def det_loss_t(p, sets, balance, box_w, obj_w, cls_w, cp=1.0, cn=0.0, ignore_obj=False, with_bbox=False, with_cls=False):
    """loss.py:138-208 / ssod_loss.py:194-288 with torch ops on p's device; sets = [certain] or [certain, uc, uc_obj, uc_cls]
    from build_targets_t."""
    dev = p[0].device
    lbox = torch.zeros(1, device=dev); lobj = torch.zeros(1, device=dev); lcls = torch.zeros(1, device=dev)
    bce = F.binary_cross_entropy_with_logits
    for l, pi in enumerate(p):
        pi = pi.float()
        nc = pi.shape[-1] - 5
        tobj = torch.zeros(pi.shape[:-1], device=dev)

        def box_term(s):
            b, a, gj, gi, tbox, anch = s[:6]
            ps = pi[b, a, gj, gi]
            pxy = ps[:, :2].sigmoid() * 2.0 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anch
            return ciou_t(torch.cat((pxy, pwh), 1), tbox), ps

        def cls_term(ps, s):
            z = torch.full_like(ps[:, 5:], cn)
            z[torch.arange(ps.shape[0], device=dev), s[6]] = cp
            return bce(ps[:, 5:], z)

        s0 = sets[0][l]
        if s0[0].shape[0]:
            iou, ps = box_term(s0)
            lbox = lbox + (1.0 - iou).mean()
            tobj[s0[0], s0[1], s0[2], s0[3]] = iou.detach().clamp(0).type(tobj.dtype)
            if nc > 1:
                lcls = lcls + cls_term(ps, s0)
        if len(sets) > 1:
            s1 = sets[1][l]
            if s1[0].shape[0]:
                tobj[s1[0], s1[1], s1[2], s1[3]] = -1.0 if ignore_obj else s1[7].type(tobj.dtype)
            if with_bbox and sets[2][l][0].shape[0]:
                iou2, _ = box_term(sets[2][l])
                lbox = lbox + (1.0 - iou2).mean()
            if with_cls and nc > 1 and sets[3][l][0].shape[0]:
                s3 = sets[3][l]
                lcls = lcls + cls_term(pi[s3[0], s3[1], s3[2], s3[3]], s3)
            valid = tobj >= 0
            lobj = lobj + bce(pi[..., 4][valid], tobj[valid]) * balance[l]
        else:
            lobj = lobj + bce(pi[..., 4], tobj) * balance[l]
    lbox = lbox * box_w; lobj = lobj * obj_w; lcls = lcls * cls_w
    return (lbox + lobj + lcls) * p[0].shape[0], (lbox.detach(), lobj.detach(), lcls.detach())


def decode_t(levels, anchors_grid, strides):
    """yolov5_head.py:66-78 with torch ops on the logits' device -> [B,P,no] fp32."""
    outs = []
    for x, anc, s in zip(levels, anchors_grid, strides):
        x = x.float()
        B, na, ny, nx, no = x.shape
        dev = x.device
        y = x.sigmoid()
        gy, gx = torch.meshgrid(torch.arange(ny, device=dev, dtype=torch.float32), torch.arange(nx, device=dev, dtype=torch.float32), indexing="ij")
        grid = torch.stack((gx, gy), 2).view(1, 1, ny, nx, 2)
        ag = (torch.as_tensor(anc, dtype=torch.float32, device=dev) * float(s)).view(1, na, 1, 1, 2)
        xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * float(s)
        wh = (y[..., 2:4] * 2.0) ** 2 * ag
        outs.append(torch.cat((xy, wh, y[..., 4:]), -1).view(B, -1, no))
    return torch.cat(outs, 1)


def nms_ssod_t(prediction, conf_thres, iou_thres, max_det=300, max_nms=30000, max_wh=7680.0):
    """general.py:887-992: Python loop per image, boolean-mask indexing (device sync each), torchvision.ops.nms."""
    import torchvision
    xc = prediction[..., 4] > conf_thres
    output = [torch.zeros((0, 8), device=prediction.device)] * prediction.shape[0]
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]
        if not x.shape[0]:
            continue
        cls_score = x[:, 5:].max(1, keepdim=True)[0]
        obj = x[:, 4:5].clone()
        x = x.clone()
        x[:, 5:] *= x[:, 4:5]
        box = torch.cat((x[:, :2] - x[:, 2:4] / 2, x[:, :2] + x[:, 2:4] / 2), 1)
        conf, j = x[:, 5:].max(1, keepdim=True)
        x = torch.cat((box, conf, j.float(), obj, cls_score), 1)[conf.view(-1) > conf_thres]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * max_wh
        i = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)
        if i.shape[0] > max_det:
            i = i[:max_det]
        output[xi] = x[i]
    return output


class EagerSSODStep:
    """One GPU-eager SSOD step (amp_dtype autocast, channels_last input).  `state_dict` has the reference's keys; a student,
    a teacher and a semi-teacher copy are kept like ModelEMA / CosineEMA do."""

    def __init__(self, state_dict, depth, neck_depth, device, anchors_grid, amp_dtype=torch.bfloat16, lr=0.01, momentum=0.937,
                 weight_decay=0.0005, batch_size=32, ema_updates=0, semi_decay=0.999, teacher_loss_weight=3.0, bn_momentum=0.03,
                 warmup=(1000, 0.1, 0.8)):
        self.dev, self.amp = torch.device(device), amp_dtype
        self.student = {k: v.detach().clone().to(self.dev) for k, v in state_dict.items()}
        self.teacher = {k: v.detach().clone() for k, v in self.student.items()}
        self.semi = {k: v.detach().clone() for k, v in self.student.items()}
        self.depth, self.neck_depth, self.anchors = depth, neck_depth, np.asarray(anchors_grid, dtype=np.float32)
        g_b, g_w, g_bn = [], [], []
        for k, v in self.student.items():
            if not v.dtype.is_floating_point or "running_" in k or k.endswith("anchors"):
                continue
            v.requires_grad_(True)
            (g_bn if k.endswith(".bn.weight") else g_b if k.endswith(".bias") else g_w).append(v)
        wd = weight_decay * batch_size * max(round(64 / batch_size), 1) / 64
        self.opt = torch.optim.SGD(g_b, lr=lr, momentum=momentum, nesterov=True)
        self.opt.add_param_group({'params': g_w, 'weight_decay': wd})
        self.opt.add_param_group({'params': g_bn})
        self.lr0, self.momentum0, self.warmup, self.ni = lr, momentum, warmup, 0
        self.ema_updates, self.semi_decay, self.tlw, self.bn_momentum = ema_updates, semi_decay, teacher_loss_weight, bn_momentum
        self.float_keys = [k for k, v in self.student.items() if v.dtype.is_floating_point]
        self.n_pseudo = 0

    def step(self, imgs, targets, u_strong, u_weak, Ms, conf_thres=0.1, iou_thres=0.65):
        H, W = u_weak.shape[2:]
        shapes = [(H // s, W // s) for s in STRIDES]
        cl = torch.channels_last
        with torch.no_grad(), torch.autocast(self.dev.type, dtype=self.amp):
            raw, _ = TrunkRef(self.teacher, self.depth, self.neck_depth).forward(u_weak.contiguous(memory_format=cl), train=False)
        pred = decode_t(raw, self.anchors, STRIDES)
        dets = nms_ssod_t(pred, conf_thres, iou_thres)
        # create_pseudo_label_online_with_gt (self_supervised_utils.py:194-245): detections go to the host, numpy float64
        rows = port.pseudo_label_rows([d.cpu().numpy() for d in dets], np.asarray(Ms.cpu() if torch.is_tensor(Ms) else Ms), H, W)
        self.n_pseudo = len(rows)
        n_img = imgs.shape[0]
        with torch.autocast(self.dev.type, dtype=self.amp):
            raw, feat = TrunkRef(self.student, self.depth, self.neck_depth, bn_momentum=self.bn_momentum).forward(
                torch.cat([imgs, u_strong], 0).contiguous(memory_format=cl), train=True)
            sup_p, un_p = [r[:n_img] for r in raw], [r[n_img:] for r in raw]
            tg = torch.as_tensor(targets, dtype=torch.float32, device=self.dev)
            sup_loss, _ = det_loss_t(sup_p, [build_targets_t(tg, self.anchors, shapes)], [4.0, 1.0, 0.4], 0.05, 0.7, 0.3)
            d0 = _domain_focal([f[:n_img] for f in feat], 0)
            d1 = _domain_focal([f[n_img:] for f in feat], 1)
            sup_loss = sup_loss + d0 * 0 + d1 * 0
            if len(rows):
                sel = [torch.as_tensor(s, device=self.dev) for s in port.select_targets(rows, [0.6] * 80, [0.1] * 80, True)]
                sets = [build_targets_t(sel[0][:, :6], self.anchors, shapes)] + [build_targets_t(s, self.anchors, shapes, with_score=True) for s in sel[1:]]
                un_loss, _ = det_loss_t(un_p, sets, [4.0, 1.0, 0.4], 0.05, 0.7, 0.3, with_bbox=True)
            else:
                un_loss = torch.zeros(1, device=self.dev)
            loss = sup_loss + un_loss * self.tlw
        loss.backward()
        if self.warmup is not None and self.ni <= self.warmup[0]:
            xi = [0, self.warmup[0]]
            for j, pg in enumerate(self.opt.param_groups):
                pg['lr'] = float(np.interp(self.ni, xi, [self.warmup[1] if j == 2 else 0.0, self.lr0]))
                pg['momentum'] = float(np.interp(self.ni, xi, [self.warmup[2], self.momentum0]))
        self.ni += 1
        self.opt.step()
        self.opt.zero_grad()
        self.ema_updates += 1
        d = 0.9999 * (1 - math.exp(-self.ema_updates / 2000))
        with torch.no_grad():       # torch_utils.py:328-338 / :405-416: a Python loop of two tensor ops per state tensor, twice
            for k in self.float_keys:
                v = self.teacher[k]
                v *= d
                v += (1.0 - d) * self.student[k].detach()
            for k in self.float_keys:
                s = self.semi[k]
                s *= self.semi_decay
                s += (1.0 - self.semi_decay) * self.teacher[k]
        return loss.detach()


def _domain_focal(feature, label):
    logits = torch.cat([f.permute(0, 2, 3, 1).reshape(-1, 2) for f in feature], 0).float()
    logp = torch.log_softmax(logits, 1)[:, label]
    return 0.5 * (-(1 - logp.exp()) ** 2 * logp).mean()
