"""`import efficientteacher_b200.bootstrap` (first line of the reference's train.py / val.py, before its trainers are
imported) rebinds the reference's hot-path symbols to the B200 mirrors -- see INTEGRATION.md section 1.  Requires the
reference checkout on sys.path (it is the host application) and libetb200.so (no CPU fallback).

The reference binds these symbols BY NAME with `from X import Y` in several places, and it loads some files twice under
two module names (`sys.path` contains both the checkout root and `models/`, so `models/loss/loss.py` exists as
`models.loss.loss` AND as `loss.loss`: models/loss/__init__.py:1, trainer/trainer.py:30).  Patching the defining module
alone would therefore leave e.g. `models.loss.build_ssod_loss` (models/loss/__init__.py:3,17-19, called from
trainer/ssod_trainer.py:261) on the original class.  apply() does three things:
  1. imports the modules that define the originals and, when importable, the ones that bind them by name
     (trainer.trainer, trainer.ssod_trainer, models.loss, val);
  2. for every patch collects ALL original objects (same __name__, defined in a module whose dotted name is a suffix of the
     defining module: `loss.loss` for `models.loss.loss`);
  3. sweeps the namespaces of every loaded module and rebinds each attribute that IS one of those objects.
Modules of the host application imported later with `from X import Y` see the patched attribute of X anyway.
"""
import importlib
import sys

from . import _lib, ema, labelmatch, loss, model, nms, pseudo_label, ssod_loss, assigner, tal

_lib.lib()  # fail loudly now if the kernels are not built


def _val_nms(original):
    """utils.general.non_max_suppression (general.py:994-1098).  The best-class variant (training path) and the val.py variant
    (`multi_label=True`, val.py:335: etb_nms_val) run on the native kernels; calls with `classes` / a-priori `labels` or CPU
    tensors keep going to the host application's own function."""
    def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                            labels=(), max_det=300):
        if classes is not None or (labels is not None and len(labels)) or not prediction.is_cuda:
            return original(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det)
        return nms.non_max_suppression(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, (), max_det)
    non_max_suppression.__wrapped__ = original
    return non_max_suppression


def _bbox_iou(original):
    """utils.metrics.bbox_iou (metrics.py:207-249): the branch the losses use (xywh, CIoU, CUDA tensors) -> etb_bbox_ciou;
    every other variant (GIoU/DIoU/plain IoU, xyxy, CPU) stays with the reference's function."""
    def bbox_iou(box1, box2, x1y1x2y2=True, GIoU=False, DIoU=False, CIoU=False, eps=1e-7):
        if (not x1y1x2y2) and CIoU and not GIoU and not DIoU and eps == 1e-7 and box1.is_cuda and box2.is_cuda \
                and box1.dim() == 2 and box2.dim() == 2 and box1.shape[0] == 4 and not (box1.requires_grad or box2.requires_grad):
            return loss.bbox_iou(box1, box2, x1y1x2y2, GIoU, DIoU, CIoU, eps)
        return original(box1, box2, x1y1x2y2, GIoU, DIoU, CIoU, eps)
    bbox_iou.__wrapped__ = original
    return bbox_iou


# (defining module, attribute, replacement | factory(original) -> replacement)
_PATCHES = [
    ("utils.torch_utils", "ModelEMA", ema.ModelEMA),
    ("utils.torch_utils", "SemiSupModelEMA", ema.SemiSupModelEMA),
    ("utils.torch_utils", "CosineEMA", ema.CosineEMA),
    ("utils.general", "non_max_suppression_ssod", nms.non_max_suppression_ssod),
    ("utils.general", "non_max_suppression", _val_nms),
    ("utils.metrics", "bbox_iou", _bbox_iou),
    ("utils.self_supervised_utils", "FairPseudoLabel", pseudo_label.FairPseudoLabel),
    ("utils.labelmatch", "LabelMatch", labelmatch.LabelMatch),
    ("models.loss.loss", "ComputeLoss", loss.ComputeLoss),
    ("models.loss.ssod.ssod_loss", "ComputeStudentMatchLoss", ssod_loss.ComputeStudentMatchLoss),
    ("models.assigner.yolo_anchor_assigner", "YOLOAnchorAssigner", assigner.YOLOAnchorAssigner),
    ("models.assigner.tal_assigner", "TaskAlignedAssigner", tal.TaskAlignedAssigner),     # the only consumer, tal_loss.py, is unimportable
    ("models.detector.yolo_ssod", "Model", model.Model),
    ("models.detector.yolo", "Model", model.SupModel),
]
_FACTORIES = (_val_nms, _bbox_iou)
# modules of the host application that bind the names above with `from X import Y` (imported here when possible so that the
# sweep reaches them; a missing optional dependency of one of them only skips that module)
_BINDERS = ["models.loss", "loss.loss", "models.assigner", "models.backbone.common", "trainer.trainer", "trainer.ssod_trainer", "val"]


def _is_original(obj, attr, mod_name):
    m = getattr(obj, "__module__", None)
    return (getattr(obj, "__name__", None) == attr and isinstance(m, str)
            and (m == mod_name or mod_name.endswith("." + m)) and not m.startswith("efficientteacher_b200"))


def apply(import_binders=True):
    """Returns the sorted list of `module.attribute` names that were rebound."""
    for mod_name, _, _ in _PATCHES:
        try:
            importlib.import_module(mod_name)
        except Exception as e:  # the host application is not on sys.path
            raise RuntimeError("efficientteacher_b200.bootstrap: cannot import reference module %s (%s)" % (mod_name, e))
    skipped = []
    if import_binders:
        for mod_name in _BINDERS:
            try:
                importlib.import_module(mod_name)
            except Exception as e:
                skipped.append((mod_name, repr(e)))
    done = []
    for mod_name, attr, repl in _PATCHES:
        defining = sys.modules[mod_name]
        originals = {}
        for m in list(sys.modules.values()):
            d = getattr(m, "__dict__", None)
            if not isinstance(d, dict):
                continue
            v = d.get(attr)
            if v is not None and _is_original(v, attr, mod_name):
                originals[id(v)] = v
        if not originals:       # already patched (apply() is idempotent)
            continue
        primary = defining.__dict__.get(attr)
        new = repl(primary if id(primary) in originals else next(iter(originals.values()))) if repl in _FACTORIES else repl
        for name, m in list(sys.modules.items()):
            d = getattr(m, "__dict__", None)
            if not isinstance(d, dict) or name.startswith("efficientteacher_b200"):
                continue
            for k, v in list(d.items()):
                if id(v) in originals:
                    d[k] = new
                    done.append(name + "." + k)
    apply.skipped = skipped
    return sorted(done)


rebound = apply()
