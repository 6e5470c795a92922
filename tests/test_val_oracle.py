"""CPU: oracle/port.process_batch (the closed-form restatement of val.py:123-145 the device kernel follows) against the golden
fixture generated from the live reference's val.process_batch (tests/golden/val_process_batch.npz)."""
import os

import numpy as np

from oracle import port

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "val_process_batch.npz")


def test_process_batch_port_matches_reference_golden():
    g = np.load(GOLD)
    iouv = np.linspace(0.5, 0.95, 10).astype(np.float32)
    import torch
    iouv = torch.linspace(0.5, 0.95, 10).numpy()
    for name in ("a", "b", "c", "d"):
        got = port.process_batch(g[name + "_det"], g[name + "_lab"], iouv)
        assert np.array_equal(got, g[name + "_correct"]), name
    assert port.process_batch(np.zeros((0, 6), np.float32), g["a_lab"], iouv).shape == (0, 10)


def test_extra_teachers_merge_port_matches_reference_golden():
    """oracle/port.merge_extra_teachers vs the live reference's create_pseudo_label_online_with_extra_teachers up to the point
    where that method raises (output_to_target_ssod on 6-column rows): tests/golden/extra_teachers.npz"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import synth
    g = np.load(os.path.join(os.path.dirname(GOLD), "extra_teachers.npz"))
    B, P = 2, 4000
    pred = synth.make_teacher_pred(21, B, P, cand_frac=0.05)
    e1 = synth.make_teacher_pred(22, B, P, cand_frac=0.04)
    e2 = synth.make_teacher_pred(23, B, P, cand_frac=0.03)
    got = port.merge_extra_teachers(pred, [e1, e2], [{3: 70, 5: 1, 7: 7}, {}], float(g["conf"]), float(g["iou"]))
    assert np.array_equal(got[0], g["out0"]) and np.array_equal(got[1], g["out1"])
