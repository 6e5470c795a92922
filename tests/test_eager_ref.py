"""CPU: the GPU-eager baseline's pieces (oracle/eager_ref.py: the reference's step restated with plain torch / torchvision
ops, timed by bench.py as `gpu_eager_baseline`) against the golden vectors of the live reference and against oracle/port.py
-- so the baseline that is timed computes the same thing as the path it is compared with."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth  # noqa: E402
from oracle import eager_ref as er  # noqa: E402
from oracle import port  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _g(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("seed,n,B,score", [(1, 128, 2, False), (4, 300, 4, True), (5, 0, 2, False)])
def test_build_targets_t_bit_exact_vs_port(seed, n, B, score):
    tg = synth.make_targets(seed, n, B) if n else np.zeros((0, 6), np.float32)
    if score:
        tg = np.concatenate([tg, np.random.RandomState(seed).uniform(0.1, 1, (len(tg), 1)).astype(np.float32)], 1)
    shapes = synth.level_shapes()
    want = port.build_targets(tg, synth.ANCHORS_GRID, shapes, with_score=score)
    got = er.build_targets_t(torch.from_numpy(tg), synth.ANCHORS_GRID, shapes, with_score=score)
    for w, (b, a, gj, gi, tbox, anch, tcls, tsc) in zip(want, got):
        assert np.array_equal(torch.stack((b, a, gj, gi), 1).numpy(), w["idx"])
        assert np.array_equal(tbox.numpy(), w["tbox"]) and np.array_equal(anch.numpy(), w["anch"]) and np.array_equal(tcls.numpy(), w["tcls"])
        if score:
            assert np.array_equal(tsc.numpy(), w["tscore"])


def test_losses_t_match_reference_golden():
    g = _g("loss_sup")
    B = int(g["B"])
    p = [torch.from_numpy(x).requires_grad_(True) for x in synth.make_head_logits(int(g["logit_seed"]), B)]
    tg = torch.from_numpy(synth.make_targets(int(g["target_seed"]), int(g["n"]), B))
    loss, (lbox, lobj, lcls) = er.det_loss_t(p, [er.build_targets_t(tg, synth.ANCHORS_GRID, synth.level_shapes())], [4.0, 1.0, 0.4], 0.05, 0.7, 0.3)
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-5)
    np.testing.assert_allclose(lbox.numpy(), g["box"], rtol=1e-5)
    np.testing.assert_allclose(lobj.numpy(), g["obj"], rtol=1e-5)
    np.testing.assert_allclose(lcls.numpy(), g["cls"], rtol=1e-5)
    g = _g("loss_ssod")
    B = int(g["B"])
    p = [torch.from_numpy(x).requires_grad_(True) for x in synth.make_head_logits(int(g["logit_seed"]), B)]
    rows = synth.make_pseudo_rows(int(g["rows_seed"]), int(g["n"]), B)
    sel = [torch.from_numpy(s) for s in port.select_targets(rows, [0.6] * 80, [0.1] * 80, with_obj=True)]
    shapes = synth.level_shapes()
    sets = [er.build_targets_t(sel[0][:, :6], synth.ANCHORS_GRID, shapes)] + [er.build_targets_t(s, synth.ANCHORS_GRID, shapes, with_score=True) for s in sel[1:]]
    loss, _ = er.det_loss_t(p, sets, [4.0, 1.0, 0.4], 0.05, 0.7, 0.3, with_bbox=True)
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-5)


def test_nms_and_decode_t_match_port():
    pred = synth.make_teacher_pred(7, 2, 25200)
    got = er.nms_ssod_t(torch.from_numpy(pred), 0.1, 0.65)
    want = port.nms_ssod(pred, 0.1, 0.65)
    for a, b in zip(got, want):
        assert np.array_equal(a.numpy(), b)
    raw = [torch.from_numpy(x) for x in synth.make_head_logits(3, 2)]
    np.testing.assert_allclose(er.decode_t(raw, synth.ANCHORS_GRID, synth.STRIDES).numpy(),
                               port.detect_decode(raw, synth.ANCHORS_GRID, synth.STRIDES).numpy(), rtol=1e-6, atol=1e-5)


def test_eager_step_matches_cpu_step_on_cpu():
    """whole eager step (fp32 on the CPU) vs oracle/step_ref.CpuSSODStep for two consecutive steps.  Not bit-identical: with 600
    pseudo-label rows many uncertain targets fall into the same (image, anchor, cell) and `tobj[b, a, gj, gi] = score`
    (ssod_loss.py:242-248) keeps whichever duplicate torch's index_put_ visits last -- the oracle port fixes "highest row wins",
    torch's TensorIterator does not promise an order -- so the objectness term differs in the 4th digit."""
    from efficientteacher_b200.config import yolov5_ssod_cfg
    from efficientteacher_b200.model import Model
    from oracle.step_ref import CpuSSODStep
    torch.manual_seed(0)
    m = Model(yolov5_ssod_cfg('l_shallow', batch_size=4, img_size=128))
    with torch.no_grad():
        for h in m.head.m:
            h.bias.view(3, -1)[:, 4] += 6.5
            h.bias.view(3, -1)[:, 5:] += 5.0
    sd = m.state_dict()
    r = np.random.RandomState(5)
    imgs = torch.from_numpy(r.rand(2, 3, 128, 128).astype(np.float32))
    uw = torch.from_numpy(r.rand(2, 3, 128, 128).astype(np.float32))
    us = uw.flip(3).contiguous()
    tg, Ms = synth.make_targets(7, 16, 2), synth.make_Ms(9, 2, 128)
    a = er.EagerSSODStep(sd, (1, 2, 3, 1), 1, "cpu", synth.ANCHORS_GRID, amp_dtype=torch.float32, batch_size=4, ema_updates=100000)
    b = CpuSSODStep(sd, (1, 2, 3, 1), 1, batch_size=4, ema_updates=100000, bn_momentum=0.03, warmup=(1000, 0.1, 0.8))
    for _ in range(2):
        la = float(a.step(imgs, tg, us, uw, Ms))
        lb, nb = b.step(imgs, tg, us, uw, Ms)
        assert a.n_pseudo == nb and nb > 0
        assert abs(la - lb) <= 3e-3 * abs(lb), (la, lb)
