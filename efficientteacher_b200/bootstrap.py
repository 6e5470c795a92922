"""`import efficientteacher_b200.bootstrap` (before the reference's trainers are imported) rebinds the reference's hot-path
symbols to the B200 mirrors -- see INTEGRATION.md section 1.  Requires the reference checkout on sys.path (it is the host
application) and a CUDA device + libetb200.so (no CPU fallback)."""
import importlib

from . import _lib, ema, labelmatch, loss, model, nms, pseudo_label, ssod_loss, assigner

_lib.lib()  # fail loudly now if the kernels are not built

_PATCHES = [
    ("utils.torch_utils", "ModelEMA", ema.ModelEMA),
    ("utils.torch_utils", "SemiSupModelEMA", ema.SemiSupModelEMA),
    ("utils.torch_utils", "CosineEMA", ema.CosineEMA),
    ("utils.general", "non_max_suppression_ssod", nms.non_max_suppression_ssod),
    ("utils.general", "non_max_suppression", "VAL_NMS"),     # see _val_nms below
    ("utils.self_supervised_utils", "FairPseudoLabel", pseudo_label.FairPseudoLabel),
    ("utils.labelmatch", "LabelMatch", labelmatch.LabelMatch),
    ("utils.self_supervised_utils", "non_max_suppression_ssod", nms.non_max_suppression_ssod),
    ("models.loss.loss", "ComputeLoss", loss.ComputeLoss),
    ("models.loss.ssod.ssod_loss", "ComputeStudentMatchLoss", ssod_loss.ComputeStudentMatchLoss),
    ("models.assigner.yolo_anchor_assigner", "YOLOAnchorAssigner", assigner.YOLOAnchorAssigner),
    ("models.detector.yolo_ssod", "Model", model.Model),
    ("models.detector.yolo", "Model", model.SupModel),
]


def _val_nms(original):
    """utils.general.non_max_suppression is also what val.py calls with multi_label=True / classes / labels -- variants that are
    not on the B200 hot path yet.  Those calls keep going to the host application's own function; the
    best-class variant (the one the training path can reach) runs on the native kernels."""
    def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                            labels=(), max_det=300):
        nc = prediction.shape[2] - 5
        if classes is not None or (multi_label and nc > 1) or (labels and len(labels)) or not prediction.is_cuda:
            return original(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det)
        return nms.non_max_suppression(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det)
    return non_max_suppression


def apply():
    done = []
    for mod_name, attr, repl in _PATCHES:
        try:
            mod = importlib.import_module(mod_name)
        except Exception as e:  # the host application is not on sys.path
            raise RuntimeError("efficientteacher_b200.bootstrap: cannot import reference module %s (%s)" % (mod_name, e))
        if repl == "VAL_NMS":
            repl = _val_nms(getattr(mod, attr))
        setattr(mod, attr, repl)
        done.append(mod_name + "." + attr)
    return done


apply()
