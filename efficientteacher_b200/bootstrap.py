"""`import efficientteacher_b200.bootstrap` (before the reference's trainers are imported) rebinds the reference's hot-path
symbols to the B200 mirrors -- see INTEGRATION.md section 1.  Requires the reference checkout on sys.path (it is the host
application) and a CUDA device + libetb200.so (no CPU fallback)."""
import importlib

from . import _lib, ema, loss, model, nms, pseudo_label, ssod_loss, assigner

_lib.lib()  # fail loudly now if the kernels are not built

_PATCHES = [
    ("utils.torch_utils", "ModelEMA", ema.ModelEMA),
    ("utils.torch_utils", "SemiSupModelEMA", ema.SemiSupModelEMA),
    ("utils.torch_utils", "CosineEMA", ema.CosineEMA),
    ("utils.general", "non_max_suppression_ssod", nms.non_max_suppression_ssod),
    ("utils.general", "non_max_suppression", nms.non_max_suppression),
    ("utils.self_supervised_utils", "FairPseudoLabel", pseudo_label.FairPseudoLabel),
    ("utils.self_supervised_utils", "non_max_suppression_ssod", nms.non_max_suppression_ssod),
    ("models.loss.loss", "ComputeLoss", loss.ComputeLoss),
    ("models.loss.ssod.ssod_loss", "ComputeStudentMatchLoss", ssod_loss.ComputeStudentMatchLoss),
    ("models.assigner.yolo_anchor_assigner", "YOLOAnchorAssigner", assigner.YOLOAnchorAssigner),
    ("models.detector.yolo_ssod", "Model", model.Model),
    ("models.detector.yolo", "Model", model.SupModel),
]


def apply():
    done = []
    for mod_name, attr, repl in _PATCHES:
        try:
            mod = importlib.import_module(mod_name)
        except Exception as e:  # the host application is not on sys.path
            raise RuntimeError("efficientteacher_b200.bootstrap: cannot import reference module %s (%s)" % (mod_name, e))
        setattr(mod, attr, repl)
        done.append(mod_name + "." + attr)
    return done


apply()
