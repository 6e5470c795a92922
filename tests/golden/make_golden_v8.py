"""Generates tests/golden/tal_*.npz and v8_*.npz from the LIVE, UNMODIFIED reference's importable YOLOv8 pieces
(`TaskAlignedAssigner`, `generate_anchors`, `YoloV8Detect`, the YOLOv8 `Model`), imported from /root/reference through
oracle/ref_harness.py.  Run in the build container only:  python tests/golden/make_golden_v8.py
Inputs are re-created from seeds by tests/synth.py, so only the outputs are stored."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

V8_YAML = 'configs/sup/public/yolov8m_coco.yaml'
HEAD_ROW_STEP = 5

# name: (seed, B, n_gt per image, img, score_pow, tiny)
TAL_CASES = {
    "a": (61, 2, [8, 8], 640, 4, 0),                  # 8 gt/img, scores ~ U(0,1)^4
    "b": (62, 3, [5, 0, 12], 640, 1, 3),              # an image without gts, padded rows, tiny boxes (top-k over zeros)
    "c": (63, 2, [40, 33], 320, 2, 0),                # crowded: many anchors claimed by several gts
    "d": (64, 1, [3], 1280, 4, 0),                    # 33600 anchors
}


def gen_tal(ns):
    from models.assigner.tal_assigner import TaskAlignedAssigner
    asg = TaskAlignedAssigner(top_k=13, num_classes=80, alpha=1.0, beta=6.0)
    for name, (seed, B, n_gt, img, sp, tiny) in TAL_CASES.items():
        d = synth.make_tal_inputs(seed, B, n_gt, img=img, score_pow=sp, tiny=tiny)
        t = {k: torch.from_numpy(v) for k, v in d.items()}
        labels, bboxes, scores, fg = asg(t["pd_scores"], t["pd_bboxes"], t["anc_points"], t["gt_labels"], t["gt_bboxes"], t["mask_gt"])
        nzi = torch.nonzero(scores)            # the score tensor is one value per foreground anchor: store it sparse
        np.savez_compressed(os.path.join(HERE, "tal_%s.npz" % name), meta=np.array([seed, B, img, sp, tiny] + list(n_gt), dtype=np.int64),
                            labels=labels.numpy(), bboxes_fg=bboxes[fg].numpy(), bboxes_bg_unique=np.unique(bboxes[~fg].numpy(), axis=0),
                            bboxes_bg_sum=np.float64(bboxes[~fg].double().sum()), fg=fg.numpy(), score_idx=nzi.numpy(),
                            score_val=scores[nzi[:, 0], nzi[:, 1], nzi[:, 2]].numpy(), score_shape=np.array(scores.shape))
        print("tal", name, "fg", int(fg.sum()), "nonzero scores", len(nzi), labels.dtype, bboxes.dtype, scores.dtype, fg.dtype)
    # n_max_boxes == 0 (tal_assigner.py:53-58): dtypes and fill values of the early return
    d = synth.make_tal_inputs(65, 2, [0, 0], img=320)
    t = {k: torch.from_numpy(v) for k, v in d.items()}
    labels, bboxes, scores, fg = asg(t["pd_scores"], t["pd_bboxes"], t["anc_points"], t["gt_labels"], t["gt_bboxes"], t["mask_gt"])
    np.savez_compressed(os.path.join(HERE, "tal_empty.npz"), labels=labels.numpy(), bboxes_absmax=bboxes.abs().max().numpy(),
                        scores_absmax=scores.abs().max().numpy(), fg=fg.numpy())
    print("tal empty", labels.dtype, labels.unique(), fg.dtype)


def gen_anchors(ns):
    from models.module.nanodet_utils import generate_anchors
    out = {}
    for img in (320, 640):
        feats = [torch.zeros(1, 1, h, w) for h, w in synth.level_shapes(img)]
        pts, st = generate_anchors(feats, [8, 16, 32], 5.0, 0.5, device='cpu', is_eval=True)
        _, pts_t, _, st_t = generate_anchors(feats, [8, 16, 32], 5.0, 0.5, device='cpu', is_eval=False)
        out["eval_pts_%d" % img], out["eval_stride_%d" % img] = pts.numpy(), st.numpy()
        out["train_pts_%d" % img], out["train_stride_%d" % img] = pts_t.numpy(), st_t.numpy()
    np.savez_compressed(os.path.join(HERE, "v8_anchors.npz"), **out)


def gen_head(ns):
    """YoloV8Detect (yolov8_head.py:117-220): eval decode from given train-layout logits.  The final 1x1 convs of the live head
    are replaced by identity-free hooks: we feed the module's own eval branch through a stub whose cv2/cv3 return our seeded logits."""
    cfg = ref_harness.make_cfg(V8_YAML, ['Model.depth_multiple', 0.33, 'Model.width_multiple', 0.25])
    torch.manual_seed(0)
    model = ns.SupModel(cfg)
    head = model.head
    out = {}
    for name, (seed, B, img) in {"a": (71, 2, 320), "b": (72, 1, 640)}.items():
        cls, reg = synth.make_v8_head_logits(seed, B, img=img, reg_max=head.reg_max)
        shapes = synth.level_shapes(img)
        # per-level NCHW maps carrying exactly these logits: cls [B,A,nc] -> [B,nc,h,w]
        offs = np.cumsum([0] + [h * w for h, w in shapes])
        cls_maps = [torch.from_numpy(cls[:, offs[i]:offs[i + 1]]).permute(0, 2, 1).reshape(B, -1, *shapes[i]).contiguous() for i in range(3)]
        reg_maps = [torch.from_numpy(reg[:, offs[i]:offs[i + 1]]).permute(0, 2, 1).reshape(B, -1, *shapes[i]).contiguous() for i in range(3)]

        class Const(torch.nn.Module):
            def __init__(self, t):
                super().__init__()
                self.t = t

            def forward(self, x):
                return self.t
        saved = head.cv2, head.cv3
        head.cv2 = torch.nn.ModuleList(Const(t) for t in reg_maps)
        head.cv3 = torch.nn.ModuleList(Const(t) for t in cls_maps)
        head.eval()
        with torch.no_grad():
            y, (feats, cls_l, reg_l) = head([torch.zeros(B, 1, *s) for s in shapes])
        head.cv2, head.cv3 = saved
        assert np.array_equal(cls_l.numpy(), cls) and np.array_equal(reg_l.numpy(), reg)      # the train layout really is [B,A,C]
        out["pred_" + name] = y.numpy()[:, ::HEAD_ROW_STEP]          # every 5th anchor row (all levels are hit)
        out["meta_" + name] = np.array([seed, B, img, head.reg_max, HEAD_ROW_STEP], dtype=np.int64)
        print("v8 head", name, y.shape)
    np.savez_compressed(os.path.join(HERE, "v8_head.npz"), **out)


def gen_model_keys(ns):
    """state_dict keys + shapes of the live YOLOv8 detector (models/detector/yolo.py with the YoloV8 backbone / neck / head) at two
    widths, and one seeded forward (train layout + eval decode) of the small one for the trunk parity test."""
    out = {}
    for tag, ov in (("m", []), ("s", ['Model.depth_multiple', 0.33, 'Model.width_multiple', 0.50])):
        cfg = ref_harness.make_cfg(V8_YAML, ov)
        torch.manual_seed(0)
        model = ns.SupModel(cfg)
        sd = model.state_dict()
        out["keys_" + tag] = np.array(list(sd.keys()))
        out["shapes_" + tag] = np.array([str(tuple(v.shape)) for v in sd.values()])
    np.savez_compressed(os.path.join(HERE, "v8_model_keys.npz"), **out)
    print("v8 keys", len(out["keys_m"]), len(out["keys_s"]))


if __name__ == "__main__":
    ns = ref_harness.load_reference()
    gen_tal(ns)
    gen_anchors(ns)
    gen_head(ns)
    gen_model_keys(ns)
