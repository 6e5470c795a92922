"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / torch-CPU fp32) of the reference's hot-path algorithms.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module,
and only as the checker.  The product (efficientteacher_b200) never imports it and has no CPU fallback.

Pinned (tests/test_oracle_vs_reference.py, run wherever /root/reference exists) against the *live, unmodified
reference* imported through oracle/ref_harness.py, and (everywhere) against the committed golden vectors in
tests/golden/*.npz that tests/golden/make_golden.py generated from that live reference.  The reference itself
ships no tests or golden vectors, and torchvision.ops.nms (0.26.0, un-vendored) is restated
from its observed semantics: stable descending score sort, IoU = inter/(a+b-inter), suppress iff IoU > thr.

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
import math

import numpy as np
import torch

F32 = np.float32


# ---------------------------------------------------------------------------------------------------------
# Detect eval decode -- models/head/yolov5_head.py:66-78
# ---------------------------------------------------------------------------------------------------------
def detect_decode(levels, anchors_grid, strides):
    """levels: list of [B,na,ny,nx,no] fp32 logits; anchors_grid [nl,na,2] (grid units); -> [B,P,no]."""
    outs = []
    for x, anc, s in zip(levels, anchors_grid, strides):
        x = torch.as_tensor(x, dtype=torch.float32)
        B, na, ny, nx, no = x.shape
        y = torch.sigmoid(x)
        gy, gx = torch.meshgrid(torch.arange(ny, dtype=torch.float32), torch.arange(nx, dtype=torch.float32), indexing="ij")
        out = y.clone()
        out[..., 0] = (y[..., 0] * 2.0 - 0.5 + gx) * float(s)
        out[..., 1] = (y[..., 1] * 2.0 - 0.5 + gy) * float(s)
        ag = (torch.as_tensor(anc, dtype=torch.float32) * float(s)).view(1, na, 1, 1, 2)
        out[..., 2:4] = (y[..., 2:4] * 2.0) ** 2 * ag
        outs.append(out.reshape(B, -1, no))
    return torch.cat(outs, 1)


# ---------------------------------------------------------------------------------------------------------
# NMS -- utils/general.py:887-992 (ssod) / 994-1098, torchvision.ops.nms semantics (call site general.py:976)
# ---------------------------------------------------------------------------------------------------------
def greedy_nms(boxes, scores, iou_thres):
    """boxes [n,4] fp32 xyxy, scores [n] fp32 -> kept indices in descending-score order (all of them)."""
    boxes = np.asarray(boxes, dtype=F32)
    order = np.argsort(-np.asarray(scores, dtype=F32), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(len(order), dtype=bool)
    keep = []
    thr = F32(iou_thres)
    for a in range(len(order)):
        if suppressed[a]:
            continue
        i = order[a]
        keep.append(i)
        rest = order[a + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(F32(0), xx2 - xx1); h = np.maximum(F32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[a + 1:] |= ovr > thr
    return np.asarray(keep, dtype=np.int64)


def nms_ssod(pred, conf_thres, iou_thres, max_det=300, max_nms=30000, max_wh=7680.0, need_cls_conf=False):
    """pred [B,P,5+nc] fp32 -> list of [k,8] fp32 arrays [x1,y1,x2,y2,conf,cls,obj,cls_score] (general.py:917-986)."""
    pred = np.asarray(pred, dtype=F32)
    thr = F32(conf_thres)
    out = []
    for x in pred:
        cand = x[:, 4] > thr                                           # :900,920
        if need_cls_conf:
            cand &= x[:, 5:].max(1) > thr                              # :1005
        x = x[cand]
        if not len(x):
            out.append(np.zeros((0, 8), F32)); continue
        cls_score = x[:, 5:].max(1)                                    # :936 (before scaling)
        cls = x[:, 5:] * x[:, 4:5]                                     # :937
        half_w, half_h = x[:, 2] / F32(2), x[:, 3] / F32(2)            # xywh2xyxy :630-637
        box = np.stack([x[:, 0] - half_w, x[:, 1] - half_h, x[:, 0] + half_w, x[:, 1] + half_h], 1)
        j = cls.argmax(1)                                              # first maximal index
        conf = cls[np.arange(len(cls)), j]
        rows = np.concatenate([box, conf[:, None], j[:, None].astype(F32), x[:, 4:5], cls_score[:, None]], 1)
        rows = rows[conf > thr]                                        # :953
        if not len(rows):
            out.append(np.zeros((0, 8), F32)); continue
        if len(rows) > max_nms:                                        # :968-969
            rows = rows[np.argsort(-rows[:, 4], kind="stable")[:max_nms]]
        c = rows[:, 5:6] * F32(max_wh)                                 # :972
        keep = greedy_nms(rows[:, :4] + c, rows[:, 4], iou_thres)[:max_det]   # :973-978
        out.append(rows[keep].astype(F32))
    return out


def nms_val(pred, conf_thres, iou_thres, multi_label=True, agnostic=False, max_det=300, max_nms=30000, max_wh=7680.0):
    """The validation-time NMS (general.py:994-1098; val.py:149-465 calls it with conf 0.001, multi_label=True):
    pred [B,P,5+nc] fp32 -> list of [k,6] fp32 arrays [x1,y1,x2,y2,conf,cls].
    multi_label: every (row, class) pair with obj*cls > conf is a detection (row-major order of torch.nonzero, :1052);
    more than max_nms detections: the max_nms best by confidence (:1071-1072); then class-offset greedy NMS, max_det."""
    pred = np.asarray(pred, dtype=F32)
    thr = F32(conf_thres)
    out = []
    for x in pred:
        nc = x.shape[1] - 5
        x = x[(x[:, 4] > thr) & (x[:, 5:].max(1) > thr)]               # :1005 candidates
        if not len(x):
            out.append(np.zeros((0, 6), F32)); continue
        cls = x[:, 5:] * x[:, 4:5]                                     # :1040 conf = obj_conf * cls_conf
        half_w, half_h = x[:, 2] / F32(2), x[:, 3] / F32(2)            # xywh2xyxy :630-637
        box = np.stack([x[:, 0] - half_w, x[:, 1] - half_h, x[:, 0] + half_w, x[:, 1] + half_h], 1)
        if multi_label and nc > 1:                                     # :1051-1053
            i, j = np.nonzero(cls > thr)
            rows = np.concatenate([box[i], cls[i, j][:, None], j[:, None].astype(F32)], 1)
        else:                                                          # :1054-1056 best class only
            j = cls.argmax(1)
            conf = cls[np.arange(len(cls)), j]
            rows = np.concatenate([box, conf[:, None], j[:, None].astype(F32)], 1)[conf > thr]
        if not len(rows):
            out.append(np.zeros((0, 6), F32)); continue
        if len(rows) > max_nms:                                        # :1071-1072
            rows = rows[np.argsort(-rows[:, 4], kind="stable")[:max_nms]]
        c = rows[:, 5:6] * F32(0.0 if agnostic else max_wh)            # :1076
        keep = greedy_nms(rows[:, :4] + c, rows[:, 4], iou_thres)[:max_det]   # :1080-1082
        out.append(rows[keep].astype(F32))
    return out


# ---------------------------------------------------------------------------------------------------------
# pseudo-label rows -- utils/plots.py:485-491 ; utils/self_supervised_utils.py:207-232, 414-454, 316-321
# ---------------------------------------------------------------------------------------------------------
def pseudo_label_rows(dets, Ms, height, width):
    """dets: list of [k,8] fp32; Ms [B,13] float64 -> float64 [N,9] rows [img,cls,cx,cy,w,h,conf,obj,cls_score]."""
    Ms = np.asarray(Ms, dtype=np.float64)
    rows_out = []
    for i, d in enumerate(dets):
        d = np.asarray(d, dtype=F32)
        if not len(d):
            continue
        # xyxy2xywh on float32 (plots.py:488-490), then widened to float64 by np.array(targets)
        cx = ((d[:, 0] + d[:, 2]) / F32(2)).astype(np.float64)
        cy = ((d[:, 1] + d[:, 3]) / F32(2)).astype(np.float64)
        w = (d[:, 2] - d[:, 0]).astype(np.float64)
        h = (d[:, 3] - d[:, 1]).astype(np.float64)
        x1, y1, x2, y2 = cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2      # xywh2xyxy, float64 (:213)
        sel = Ms[Ms[:, 0] == i][0]
        M = sel[1:10].reshape(3, 3); s = float(sel[10]); ud = int(sel[11]); lr = int(sel[12])
        px = np.stack([x1, x2, x1, x2], 1); py = np.stack([y1, y2, y2, y1], 1)  # x1y1,x2y2,x1y2,x2y1 (:438)
        ox = px * M[0, 0] + py * M[0, 1] + M[0, 2]
        oy = px * M[1, 0] + py * M[1, 1] + M[1, 2]
        nx1 = ox.min(1).clip(0, width); nx2 = ox.max(1).clip(0, width)
        ny1 = oy.min(1).clip(0, height); ny2 = oy.max(1).clip(0, height)
        w1, h1 = x2 * s - x1 * s, y2 * s - y1 * s                              # box_candidates (:316-321)
        w2, h2 = nx2 - nx1, ny2 - ny1
        ar = np.maximum(w2 / (h2 + 1e-16), h2 / (w2 + 1e-16))
        ok = (w2 > 2) & (h2 > 2) & (w2 * h2 / (w1 * h1 + 1e-16) > 0.10) & (ar < 20)
        ocx, ocy, ow, oh = (nx1 + nx2) / 2, (ny1 + ny2) / 2, nx2 - nx1, ny2 - ny1
        ocx, ow, ocy, oh = ocx / width, ow / width, ocy / height, oh / height   # :225-226
        if ud == 1:
            ocy = 1 - ocy
        if lr == 1:
            ocx = 1 - ocx
        r = np.stack([np.full(len(d), float(i)), d[:, 5].astype(np.float64), ocx, ocy, ow, oh,
                      d[:, 4].astype(np.float64), d[:, 6].astype(np.float64), d[:, 7].astype(np.float64)], 1)
        rows_out.append(r[ok])
    if not rows_out:
        return np.zeros((0, 9))
    return np.concatenate(rows_out, 0)


# ---------------------------------------------------------------------------------------------------------
# select_targets -- models/loss/ssod/ssod_loss.py:130-192
# ---------------------------------------------------------------------------------------------------------
def select_targets(rows, thr_high, thr_low, with_obj=True):
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 9)
    hi = np.asarray(thr_high, dtype=np.float64); lo = np.asarray(thr_low, dtype=np.float64)
    c = rows[:, 1].astype(np.int64)
    rel = rows[:, 6] >= hi[c]
    unc = (~rel) & (rows[:, 6] >= lo[c])
    uc_rows = np.concatenate([rows[:, :6], rows[:, 7:8] if with_obj else rows[:, 6:7]], 1)
    out = [rows[rel][:, :7], uc_rows[unc]]
    if with_obj:
        out.append(uc_rows[unc & (rows[:, 7] >= 0.99)])
        out.append(uc_rows[unc & (rows[:, 8] >= 0.99)])
    else:
        out += [np.zeros((0, 7)), np.zeros((0, 7))]
    return [o.astype(F32) for o in out]


# ---------------------------------------------------------------------------------------------------------
# build_targets / build_uc_targets_aug -- models/assigner/yolo_anchor_assigner.py:319-372, 640-697
# ---------------------------------------------------------------------------------------------------------
def build_targets(targets, anchors_grid, shapes, anchor_t=4.0, with_score=False):
    """targets [nt,6|7] fp32; shapes: list of (ny,nx).  Returns per level dict(idx[n,4] int64 (b,a,gj,gi),
    tbox[n,4], anch[n,2], tcls[n] int64, tscore[n])."""
    ts = 7 if with_score else 6
    t_in = np.asarray(targets, dtype=F32).reshape(-1, targets.shape[1] if hasattr(targets, "shape") and len(targets.shape) > 1 else ts)[:, :ts]
    nt = len(t_in)
    res = []
    offs = np.array([[0, 0], [0.5, 0], [0, 0.5], [-0.5, 0], [0, -0.5]], dtype=F32)
    for l, (ny, nx) in enumerate(shapes):
        A = np.asarray(anchors_grid[l], dtype=F32).reshape(3, 2)
        empty = dict(idx=np.zeros((0, 4), np.int64), tbox=np.zeros((0, 4), F32), anch=np.zeros((0, 2), F32),
                     tcls=np.zeros((0,), np.int64), tscore=np.zeros((0,), F32))
        if nt == 0:
            res.append(empty); continue
        gx = t_in[:, 2] * F32(nx); gy = t_in[:, 3] * F32(ny)
        gw = t_in[:, 4] * F32(nx); gh = t_in[:, 5] * F32(ny)
        rows = []
        for a in range(3):                                     # anchor-major, target order (:342-346)
            rw, rh = gw / A[a, 0], gh / A[a, 1]
            m = np.maximum(np.maximum(rw, F32(1) / rw), np.maximum(rh, F32(1) / rh)) < F32(anchor_t)
            k = np.nonzero(m)[0]
            rows.append(np.stack([np.full(len(k), a), k], 1))
        ak = np.concatenate(rows, 0)
        a_i, k_i = ak[:, 0], ak[:, 1]
        mgx, mgy = gx[k_i], gy[k_i]
        ix, iy = F32(nx) - mgx, F32(ny) - mgy                  # gxi (:349)
        flags = [np.ones(len(k_i), bool),
                 (np.mod(mgx, F32(1)) < F32(0.5)) & (mgx > 1), (np.mod(mgy, F32(1)) < F32(0.5)) & (mgy > 1),
                 (np.mod(ix, F32(1)) < F32(0.5)) & (ix > 1), (np.mod(iy, F32(1)) < F32(0.5)) & (iy > 1)]
        idx, tbox, anch, tcls, tsc = [], [], [], [], []
        for o in range(5):                                     # offset-major (:352-354)
            f = flags[o]
            kk, aa = k_i[f], a_i[f]
            ox, oy = mgx[f] - offs[o, 0], mgy[f] - offs[o, 1]
            gi0 = np.trunc(ox).astype(np.int64); gj0 = np.trunc(oy).astype(np.int64)   # .long()
            gi = np.clip(gi0, 0, nx - 1); gj = np.clip(gj0, 0, ny - 1)
            idx.append(np.stack([t_in[kk, 0].astype(np.int64), aa, gj, gi], 1))
            tbox.append(np.stack([mgx[f] - gi0.astype(F32), mgy[f] - gj0.astype(F32), gw[kk], gh[kk]], 1))
            anch.append(A[aa])
            tcls.append(t_in[kk, 1].astype(np.int64))
            tsc.append(t_in[kk, 6] if with_score else np.zeros(len(kk), F32))
        res.append(dict(idx=np.concatenate(idx), tbox=np.concatenate(tbox).astype(F32), anch=np.concatenate(anch).astype(F32),
                        tcls=np.concatenate(tcls), tscore=np.concatenate(tsc).astype(F32)))
    return res


# ---------------------------------------------------------------------------------------------------------
# CIoU + losses (torch CPU fp32, differentiable) -- utils/metrics.py:207-249 ; models/loss/loss.py:138-208 ;
# models/loss/ssod/ssod_loss.py:194-288
# ---------------------------------------------------------------------------------------------------------
def ciou(b1, b2, eps=1e-7):
    """b1,b2 [n,4] centre/size torch tensors -> [n]."""
    b1x1, b1x2 = b1[:, 0] - b1[:, 2] / 2, b1[:, 0] + b1[:, 2] / 2
    b1y1, b1y2 = b1[:, 1] - b1[:, 3] / 2, b1[:, 1] + b1[:, 3] / 2
    b2x1, b2x2 = b2[:, 0] - b2[:, 2] / 2, b2[:, 0] + b2[:, 2] / 2
    b2y1, b2y2 = b2[:, 1] - b2[:, 3] / 2, b2[:, 1] + b2[:, 3] / 2
    inter = (torch.min(b1x2, b2x2) - torch.max(b1x1, b2x1)).clamp(0) * (torch.min(b1y2, b2y2) - torch.max(b1y1, b2y1)).clamp(0)
    w1, h1 = b1x2 - b1x1, b1y2 - b1y1 + eps
    w2, h2 = b2x2 - b2x1, b2y2 - b2y1 + eps
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.max(b1x2, b2x2) - torch.min(b1x1, b2x1)
    ch = torch.max(b1y2, b2y2) - torch.min(b1y1, b2y1)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((b2x1 + b2x2 - b1x1 - b1x2) ** 2 + (b2y1 + b2y2 - b1y1 - b1y2) ** 2) / 4
    v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)) ** 2
    alpha = (v / (v - iou + (1 + eps))).detach()
    return iou - (rho2 / c2 + v * alpha)


def _bce(x, z):
    return torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs()))


def _t(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def det_loss(p, sets, balance, box_w, obj_w, cls_w, cp=1.0, cn=0.0, ignore_obj=False, with_bbox=False, with_cls=False):
    """p: list of [B,na,ny,nx,no] torch fp32 (may require grad).  sets: [certain] or [certain, uc, uc_obj, uc_cls],
    each the per-level list build_targets returns.  -> (loss*B [1], (lbox, lobj, lcls))."""
    lbox = torch.zeros(1); lobj = torch.zeros(1); lcls = torch.zeros(1)
    for l, pi in enumerate(p):
        nc = pi.shape[-1] - 5
        tobj = torch.zeros(pi.shape[:-1])

        def gather(s):
            idx = _t(s["idx"], torch.int64)
            return pi[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]], idx

        def box_term(s):
            ps, idx = gather(s)
            pxy = ps[:, :2].sigmoid() * 2.0 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * _t(s["anch"])
            return ciou(torch.cat([pxy, pwh], 1), _t(s["tbox"])), ps, idx

        def cls_term(ps, s):
            z = torch.full_like(ps[:, 5:], cn)
            z[torch.arange(len(ps)), _t(s["tcls"], torch.int64)] = cp
            return _bce(ps[:, 5:], z).mean()

        s0 = sets[0][l]
        if len(s0["idx"]):
            iou, ps, idx = box_term(s0)
            lbox = lbox + (1.0 - iou).mean()
            v = iou.detach().clamp(0)
            for r in range(len(idx)):                           # last row wins (CPU index_put_ semantics)
                tobj[idx[r, 0], idx[r, 1], idx[r, 2], idx[r, 3]] = v[r]
            if nc > 1:
                lcls = lcls + cls_term(ps, s0)
        if len(sets) > 1:
            s1 = sets[1][l]
            idx1 = _t(s1["idx"], torch.int64)
            sc = _t(s1["tscore"])
            for r in range(len(idx1)):
                tobj[idx1[r, 0], idx1[r, 1], idx1[r, 2], idx1[r, 3]] = -1.0 if ignore_obj else sc[r]
            if with_bbox and len(sets[2][l]["idx"]):
                iou2, _, _ = box_term(sets[2][l])
                lbox = lbox + (1.0 - iou2).mean()
            if with_cls and nc > 1 and len(sets[3][l]["idx"]):
                ps3, _ = gather(sets[3][l])
                lcls = lcls + cls_term(ps3, sets[3][l])
        valid = tobj >= 0
        lobj = lobj + _bce(pi[..., 4][valid], tobj[valid]).mean() * balance[l]
    lbox = lbox * box_w; lobj = lobj * obj_w; lcls = lcls * cls_w
    B = p[0].shape[0]
    return (lbox + lobj + lcls) * B, (lbox, lobj, lcls)


# ---------------------------------------------------------------------------------------------------------
# EMA -- utils/torch_utils.py:328-338
# ---------------------------------------------------------------------------------------------------------
def ema_update(v, m, d):
    """v <- v*d ; v += (1-d)*m  with the python scalars rounded to fp32.  numpy fp32 arrays."""
    v = np.asarray(v, dtype=F32); m = np.asarray(m, dtype=F32)
    return (v * F32(d)) + (F32(1.0 - d) * m)


# ---------------------------------------------------------------------------------------------------------
# validation matching -- val.py:123-145 (process_batch) ; utils/metrics.py:252-273 (box_iou)
# ---------------------------------------------------------------------------------------------------------
def process_batch(detections, labels, iouv):
    """detections [N,6] (x1,y1,x2,y2,conf,cls), labels [M,5] (cls,x1,y1,x2,y2), iouv [T] -> correct [N,T] bool.
    Restated closed form of the reference's sort / np.unique sequence (see csrc/val.cu): per detection the best class-matching
    label (later label on IoU ties); a detection is correct at threshold t iff its IoU >= t and no EARLIER detection with the
    same best label also reaches t."""
    det = np.asarray(detections, dtype=F32); lab = np.asarray(labels, dtype=F32); thr = np.asarray(iouv, dtype=F32)
    N, M = len(det), len(lab)
    correct = np.zeros((N, len(thr)), bool)
    if N == 0 or M == 0:
        return correct
    a1 = (lab[:, 3] - lab[:, 1]) * (lab[:, 4] - lab[:, 2])
    a2 = (det[:, 2] - det[:, 0]) * (det[:, 3] - det[:, 1])
    w = np.maximum(np.minimum(lab[:, None, 3], det[None, :, 2]) - np.maximum(lab[:, None, 1], det[None, :, 0]), F32(0))
    h = np.maximum(np.minimum(lab[:, None, 4], det[None, :, 3]) - np.maximum(lab[:, None, 2], det[None, :, 1]), F32(0))
    inter = (w * h).astype(F32)
    iou = inter / (a1[:, None] + a2[None, :] - inter)
    same = lab[:, 0:1] == det[None, :, 5]
    iou_m = np.where(same, iou, F32(-1))
    best_l = (M - 1) - np.argmax(iou_m[::-1], axis=0)            # last maximal label
    best = iou_m[best_l, np.arange(N)]
    for i, t in enumerate(thr):
        taken = set()
        for d in range(N):
            if best[d] >= t and best[d] >= 0:
                if best_l[d] not in taken:
                    correct[d, i] = True
                    taken.add(best_l[d])
    return correct


# ---------------------------------------------------------------------------------------------------------
# extra-teachers merge -- utils/self_supervised_utils.py:256-274 (the part of the method that can execute)
# ---------------------------------------------------------------------------------------------------------
def merge_extra_teachers(pred, extra_preds, class_maps, conf_thres, iou_thres):
    """pred / extra_preds[t]: [B,P,5+nc] fp32; class_maps[t]: {origin_cls: new_cls}.  -> list of [k,6] fp32 arrays."""
    cur = nms_val(pred, conf_thres, iou_thres, multi_label=False)
    for t, tp in enumerate(extra_preds):
        td = nms_val(tp, conf_thres, iou_thres, multi_label=False)
        for i in range(len(cur)):
            d = td[i].copy()
            for r in d:                                            # :264-268
                if int(r[5]) in class_maps[t]:
                    r[5] = F32(class_maps[t][int(r[5])])
            x = np.concatenate([cur[i], d], 0)                     # :270
            keep = greedy_nms(x[:, :4], x[:, 4], iou_thres)        # :272-274 (c = 0: class-agnostic)
            cur[i] = x[keep]
    return cur
