"""CPU: `import efficientteacher_b200.bootstrap` with the live reference on sys.path rebinds every name the reference's
trainers construct on the hot path (ssod_trainer.py:69,104,123-131,261; trainer.py:320-321) -- including the
by-name aliases (`models.loss.build_ssod_loss` resolves `ComputeStudentMatchLoss` from models/loss/__init__.py:3) and the
twice-loaded modules (`loss.loss` vs `models.loss.loss`).  Runs in a subprocess: the patch is process-wide.
Needs /root/reference (absent on the GPU box -> skipped there)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ETB_REFERENCE_ROOT", "/root/reference")

SCRIPT = textwrap.dedent('''
    import sys
    sys.path.insert(0, %r)
    from oracle import ref_harness                  # test infrastructure: import shims for the unmodified reference
    ref_harness.load_reference()
    import efficientteacher_b200.bootstrap as bs
    assert not bs.apply.skipped, bs.apply.skipped
    from efficientteacher_b200 import assigner, ema, labelmatch, loss, model, nms, pseudo_label, ssod_loss
    import trainer.trainer as T
    import trainer.ssod_trainer as S
    import val as V
    ML = sys.modules["models.loss"]
    # trainer/ssod_trainer.py:261 -> models.loss.build_ssod_loss -> the name bound in models/loss/__init__.py:3
    assert ML.ComputeStudentMatchLoss is ssod_loss.ComputeStudentMatchLoss
    assert S.build_ssod_loss is ML.build_ssod_loss and S.build_ssod_loss.__globals__["ComputeStudentMatchLoss"] is ssod_loss.ComputeStudentMatchLoss
    assert ML.build_loss.__globals__["ComputeLoss"] is loss.ComputeLoss
    assert sys.modules["loss.loss"].ComputeLoss is loss.ComputeLoss and sys.modules["models.loss.loss"].ComputeLoss is loss.ComputeLoss
    # trainer/trainer.py:320-321, :133,146, :157
    assert T.ComputeLoss is loss.ComputeLoss and T.Model is model.SupModel and T.ModelEMA is ema.ModelEMA
    # trainer/ssod_trainer.py:69,104,114,123-131
    assert S.Model is model.Model and S.FairPseudoLabel is pseudo_label.FairPseudoLabel and S.LabelMatch is labelmatch.LabelMatch
    assert S.ModelEMA is ema.ModelEMA and S.SemiSupModelEMA is ema.SemiSupModelEMA and S.CosineEMA is ema.CosineEMA
    # pseudo-label creators call the NMS through their own module globals (self_supervised_utils.py:27, labelmatch.py:26)
    import utils.self_supervised_utils as U
    assert U.non_max_suppression_ssod is nms.non_max_suppression_ssod
    # val.py:335 and the wrapped variants
    import utils.general as G, utils.metrics as MT
    assert V.non_max_suppression is G.non_max_suppression and G.non_max_suppression.__module__ == "efficientteacher_b200.bootstrap"
    assert MT.bbox_iou.__module__ == "efficientteacher_b200.bootstrap"
    import torch
    b1, b2 = torch.rand(4, 7), torch.rand(7, 4)
    assert torch.equal(MT.bbox_iou(b1, b2, x1y1x2y2=False, CIoU=True), MT.bbox_iou.__wrapped__(b1, b2, x1y1x2y2=False, CIoU=True))  # CPU -> reference
    assert sys.modules["models.assigner"].YOLOAnchorAssigner is assigner.YOLOAnchorAssigner
    from efficientteacher_b200 import tal
    assert sys.modules["models.assigner.tal_assigner"].TaskAlignedAssigner is tal.TaskAlignedAssigner
    assert bs.apply() == []          # idempotent
    # nothing of the reference's own hot-path classes is left reachable from the trainers' namespaces
    for mod in (T, S):
        for k, v in vars(mod).items():
            m = getattr(v, "__module__", "") or ""
            assert not (k in ("ComputeLoss", "Model", "ModelEMA", "FairPseudoLabel", "LabelMatch", "CosineEMA", "SemiSupModelEMA")
                        and not m.startswith("efficientteacher_b200")), (mod.__name__, k, m)
    print("BOOTSTRAP_OK", len(bs.rebound))
''')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "trainer")), reason="reference checkout not present")
def test_bootstrap_rebinds_every_hot_path_symbol():
    env = dict(os.environ, WANDB_MODE="disabled", PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "BOOTSTRAP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
