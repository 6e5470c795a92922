"""FairPseudoLabel with the reference's surface (utils/self_supervised_utils.py:54-245).

create_pseudo_label_online_with_gt keeps the reference's return contract -- (CPU float64 [N,9] tensor, invalid
flag) because trainer/ssod_trainer.py:618,640,645,663-669 consumes it on the host -- but the whole chain
(candidate filter -> NMS -> xyxy2xywh -> affine warp to the strong-aug frame -> box_candidates -> normalise ->
flips) runs on the device in one batch of launches; the single D2H copy is the final [N,9] rows.  A
device-resident twin (`last_rows_dev`, `last_count_dev`) is cached for the fused SSOD loss so the step itself
never waits on that copy.
"""
import torch

from . import _lib
from .nms import _run


class FairPseudoLabel:
    def __init__(self, cfg):
        self.nms_conf_thres = cfg.SSOD.nms_conf_thres
        self.nms_iou_thres = cfg.SSOD.nms_iou_thres
        self.debug = cfg.SSOD.debug
        self.multi_label = cfg.SSOD.multi_label
        self.names = cfg.Dataset.names
        self.num_points = cfg.Dataset.np
        if self.multi_label or self.num_points:
            raise NotImplementedError("SSOD.multi_label / keypoints are not on the B200 hot path")
        self.last_rows_dev = None
        self.last_count_dev = None
        self.last_det = None

    def create_pseudo_label_device(self, out, M_s, height, width):
        """Device-only variant: returns (rows float64 [B*300,9] CUDA, count int32[1] CUDA); no host sync."""
        Ms = M_s.to(device=out.device, dtype=torch.float64, non_blocking=True).contiguous()
        assert Ms.shape == (out.shape[0], 13), "M_s must be [B,13] = [img, M(9), s, ud, lr]"
        det, det_cnt, rows, cnt = _run(out, self.nms_conf_thres, self.nms_iou_thres, False, 300, False, Ms=Ms,
                                       img_hw=(height, width))
        self.last_rows_dev, self.last_count_dev, self.last_det = rows, cnt, (det, det_cnt)
        return rows, cnt

    def create_pseudo_label_online_with_gt(self, out, target_imgs, M_s, target_imgs_ori, gt=None, RANK=-2):
        n_img, _, height, width = target_imgs.shape
        rows, cnt = self.create_pseudo_label_device(out.detach(), M_s, height, width)
        n = int(cnt.item())
        if n == 0:
            return [], True
        return rows[:n].cpu(), False
