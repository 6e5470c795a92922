"""TEST INFRASTRUCTURE ONLY -- live import of the *unmodified* reference.

Only tests/, tests/golden/make_golden.py and bench.py's cpu/reference leg may import this module.
It needs /root/reference (present in the build container, absent on the GPU box), so every caller
must gate on `reference_available()`.

Nothing from the reference is copied: we put it on sys.path and apply four harness-side shims
(matplotlib/seaborn stubs, YOLOV5_CONFIG_DIR, clamp_ dtype cast, .cuda() no-op on CPU boxes).
"""
import importlib
import os
import sys
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("ETB_REFERENCE_ROOT", "/root/reference")
_loaded = False


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models")) and os.path.isdir(os.path.join(REF_ROOT, "utils"))


def load_reference():
    """Import-shim the reference; idempotent. Returns a namespace of the hot-path symbols."""
    global _loaded
    if not reference_available():
        raise RuntimeError("reference tree not available at %s" % REF_ROOT)
    import numpy as np
    import torch
    if not _loaded:
        sys.dont_write_bytecode = True  # the reference dir is read-only
        os.environ.setdefault("YOLOV5_CONFIG_DIR", os.path.join(REF_ROOT, "utils"))
        for n in ("matplotlib", "matplotlib.pyplot", "seaborn", "thop"):
            try:
                importlib.import_module(n)
            except Exception:
                sys.modules[n] = MagicMock()
        for p in (os.path.join(REF_ROOT, "models"), REF_ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        _c = torch.Tensor.clamp_

        def clamp_(self, min=None, max=None):  # yolo_anchor_assigner.py:367,691 on torch>=1.12
            if not self.is_floating_point():
                if isinstance(min, torch.Tensor):
                    min = min.to(self.dtype)
                if isinstance(max, torch.Tensor):
                    max = max.to(self.dtype)
            return _c(self, min, max)

        torch.Tensor.clamp_ = clamp_
        if not torch.cuda.is_available():  # models/loss/loss.py:392,418 hard-code .cuda()
            torch.Tensor.cuda = lambda s, *a, **k: s
            torch.nn.Module.cuda = lambda s, *a, **k: s
        if not hasattr(np, "int"):
            np.int = int  # utils/general.py:516,531
        _loaded = True

    from types import SimpleNamespace
    ns = SimpleNamespace()
    from configs.defaults import get_cfg
    from models.detector.yolo_ssod import Model as SSODModel
    from models.detector.yolo import Model as SupModel
    from models.loss.loss import ComputeLoss, DomainLoss, TargetLoss
    from models.loss.ssod.ssod_loss import ComputeStudentMatchLoss
    from models.assigner.yolo_anchor_assigner import YOLOAnchorAssigner
    from utils.general import non_max_suppression_ssod, non_max_suppression, xywh2xyxy, xyxy2xywh
    from utils.metrics import bbox_iou, box_iou
    from utils.plots import output_to_target_ssod
    from utils.self_supervised_utils import FairPseudoLabel
    from utils.torch_utils import ModelEMA, SemiSupModelEMA, CosineEMA
    ns.__dict__.update(locals())
    ns.REF_ROOT = REF_ROOT
    return ns


def make_cfg(yaml_rel, overrides=()):
    ns = load_reference()
    cfg = ns.get_cfg()
    cfg.merge_from_file(os.path.join(REF_ROOT, yaml_rel))
    if overrides:
        cfg.merge_from_list(list(overrides))
    return cfg
