"""The per-batch work of the reference's validation pass (val.py:279-372) on the device: half-precision forward (the native
engine computes in bf16), multi-label NMS (etb_nms_val), rescale to the native image space, and the true-positive matching
process_batch (etb_val_process_batch) for all images of the batch in one launch.  The epoch-level bookkeeping
(ap_per_class, confusion matrix, COCO json) stays with the host application."""
import torch

from . import _lib
from . import nms as etb_nms


def box_iou(box1, box2):
    """utils/metrics.py:252-273 for callers that want the matrix itself (plain torch ops; not on the hot path)"""
    area1 = (box1[:, 2] - box1[:, 0]) * (box1[:, 3] - box1[:, 1])
    area2 = (box2[:, 2] - box2[:, 0]) * (box2[:, 3] - box2[:, 1])
    inter = (torch.min(box1[:, None, 2:], box2[:, 2:]) - torch.max(box1[:, None, :2], box2[:, :2])).clamp(0).prod(2)
    return inter / (area1[:, None] + area2 - inter)


def process_batch_batched(det, det_cnt, labels, iouv):
    """det [B,max_det,>=6] fp32 (x1,y1,x2,y2,conf,cls) and labels [nt,6] (img,cls,x1,y1,x2,y2) in the same coordinate space;
    det_cnt [B] int32 or None; iouv [T] -> correct [B,max_det,T] bool (rows >= det_cnt are False)."""
    _lib.require_cuda(det, iouv)
    det = det.float().contiguous()
    labels = labels.to(det.device).float().contiguous()
    iouv = iouv.to(det.device).float().contiguous()
    B, max_det, ld = det.shape
    T = iouv.numel()
    correct = torch.empty((B, max_det, T), dtype=torch.uint8, device=det.device)
    overflow = torch.zeros(1, dtype=torch.int32, device=det.device)
    _lib.check(_lib.lib().etb_val_process_batch(_lib.ptr(det), _lib.ptr(det_cnt), B, max_det, ld, _lib.ptr(labels), int(labels.shape[0]),
                                                _lib.ptr(iouv), T, _lib.ptr(correct), _lib.ptr(overflow), _lib.stream_ptr()),
               "etb_val_process_batch")
    return correct.bool()


def process_batch(detections, labels, iouv):
    """val.py:123-145: detections [N,6] (x1,y1,x2,y2,conf,cls), labels [M,5] (cls,x1,y1,x2,y2) -> correct [N,len(iouv)] bool"""
    n = detections.shape[0]
    if n == 0:
        return torch.zeros((0, iouv.numel()), dtype=torch.bool, device=detections.device)
    lab = torch.cat([torch.zeros((labels.shape[0], 1), device=labels.device, dtype=labels.dtype), labels], 1)
    return process_batch_batched(detections[None, :, :6], None, lab, iouv)[0]


def scale_coords_(img1_shape, coords, img0_shape, ratio_pad=None):
    """utils/general.py:702-715 (in place, xyxy in columns 0..3)"""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    coords[:, 0].clamp_(0, img0_shape[1]); coords[:, 1].clamp_(0, img0_shape[0])
    coords[:, 2].clamp_(0, img0_shape[1]); coords[:, 3].clamp_(0, img0_shape[0])
    return coords


@torch.no_grad()
def val_batch(model, img, targets, shapes, conf_thres=0.001, iou_thres=0.6, iouv=None, single_cls=False):
    """One batch of val.run (val.py:300-372, detection branch): img [B,3,H,W] (uint8 or fp32 in [0,1]) -> forward (native
    engine) -> non_max_suppression(multi_label=True) -> native-space predictions -> process_batch.  targets [nt,6] (img, cls,
    xywh normalised); shapes[i] = (shape0, (ratio, pad)) like the reference's loader.  Returns a list of (correct [n,T] bool,
    conf [n], pcls [n], tcls list) per image -- the tuples val.py appends to `stats`."""
    if iouv is None:
        iouv = torch.linspace(0.5, 0.95, 10, device=img.device)
    B, _, H, W = img.shape
    out = model(img)
    pred = out[0] if isinstance(out, tuple) else out
    pred = pred[0] if isinstance(pred, tuple) else pred            # SSOD model: ((pred, raw), features)
    dets = etb_nms.non_max_suppression(pred, conf_thres, iou_thres, multi_label=True, agnostic=single_cls)
    tg = targets.to(img.device).float().clone()
    tg[:, 2:6] *= torch.tensor([W, H, W, H], device=img.device, dtype=torch.float32)
    max_det = max(max((d.shape[0] for d in dets), default=0), 1)
    det_pad = torch.zeros((B, max_det, 6), dtype=torch.float32, device=img.device)
    cnt = torch.zeros(B, dtype=torch.int32, device=img.device)
    labs = []
    for si, p in enumerate(dets):
        predn = p.clone()
        if single_cls:
            predn[:, 5] = 0
        scale_coords_((H, W), predn[:, :4], shapes[si][0], shapes[si][1])
        det_pad[si, :predn.shape[0]] = predn
        cnt[si] = predn.shape[0]
        l = tg[tg[:, 0] == si, 1:]
        if l.shape[0]:
            tb = torch.cat((l[:, 1:3] - l[:, 3:5] / 2, l[:, 1:3] + l[:, 3:5] / 2), 1)       # xywh2xyxy
            scale_coords_((H, W), tb, shapes[si][0], shapes[si][1])
            labs.append(torch.cat((torch.full((l.shape[0], 1), float(si), device=img.device), l[:, 0:1], tb), 1))
    lab = torch.cat(labs, 0) if labs else torch.zeros((0, 6), device=img.device)
    correct = process_batch_batched(det_pad, cnt, lab, iouv)
    stats = []
    for si, p in enumerate(dets):
        n = p.shape[0]
        tcls = tg[tg[:, 0] == si, 1].tolist()
        stats.append((correct[si, :n], det_pad[si, :n, 4], det_pad[si, :n, 5], tcls))
    return stats
