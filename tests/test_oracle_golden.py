"""CPU: the oracle restatement (oracle/port.py) against the golden vectors produced by the live, unmodified
reference (tests/golden/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare the
CUDA kernels with the oracle and with the same golden vectors."""
import numpy as np
import pytest
import torch

import synth
from oracle import port


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_build_targets_bit_exact(golden, name):
    g = golden("assign_" + name)
    B, n = int(g["B"]), int(g["n"])
    t = synth.make_targets(int(g["seed"]), n, B)
    sc = np.random.RandomState(int(g["score_seed"])).uniform(0.1, 1, (n, 1)).astype(np.float32)
    for pref, tt, ws in (("bt_", t, False), ("uc_", np.concatenate([t, sc], 1), True)):
        res = port.build_targets(tt, synth.ANCHORS_GRID, synth.level_shapes(), 4.0, with_score=ws)
        for l in range(3):
            assert np.array_equal(res[l]["idx"], g[f"{pref}idx{l}"]), (pref, l)
            assert np.array_equal(res[l]["tcls"], g[f"{pref}tcls{l}"])
            assert np.array_equal(res[l]["tbox"], g[f"{pref}tbox{l}"])
            assert np.array_equal(res[l]["anch"], g[f"{pref}anch{l}"])
            if ws:
                assert np.array_equal(res[l]["tscore"], g[f"{pref}tscore{l}"])


def test_build_targets_empty():
    res = port.build_targets(np.zeros((0, 6), np.float32), synth.ANCHORS_GRID, synth.level_shapes())
    assert all(len(r["idx"]) == 0 for r in res)


@pytest.mark.parametrize("name", ["a", "dense", "cap", "hi"])
def test_nms_and_pseudo_rows(golden, name):
    g = golden("nms_" + name)
    B, P = int(g["B"]), int(g["P"])
    pred = synth.make_teacher_pred(int(g["seed"]), B, P, cand_frac=float(g["frac"]))
    if name == "a":
        pred[2, :, 4] = 0.01
    dets = port.nms_ssod(pred, float(g["conf_thres"]), float(g["iou_thres"]))
    for b in range(B):
        assert np.array_equal(dets[b], g[f"det{b}"]), (name, b)      # keep-set, order and values bit-exact
    rows = port.pseudo_label_rows(dets, g["Ms"], 640, 640)
    assert rows.shape == g["rows"].shape
    assert np.array_equal(rows[:, :2], g["rows"][:, :2])
    np.testing.assert_allclose(rows, g["rows"], rtol=1e-12, atol=1e-12)
    if f"val0" in g.files:
        d6 = port.nms_ssod(pred, 0.25, 0.45, need_cls_conf=True)
        for b in range(B):
            assert np.array_equal(d6[b][:, :6], g[f"val{b}"])


def test_select_targets(golden):
    g = golden("select")
    rows = synth.make_pseudo_rows(int(g["seed"]), int(g["n"]), int(g["B"]))
    sel = port.select_targets(rows, g["high"], g["low"], with_obj=True)
    for i in range(4):
        assert np.array_equal(sel[i], g[f"s{i}"]), i


def test_ciou(golden):
    g = golden("ciou")
    c = port.ciou(torch.from_numpy(g["b1"]), torch.from_numpy(g["b2"])).numpy()
    np.testing.assert_allclose(c, g["ciou"], rtol=1e-6, atol=1e-6)


def test_decode(golden):
    g = golden("decode")
    pred = port.detect_decode([g["raw0"], g["raw1"], g["raw2"]], synth.ANCHORS_GRID, synth.STRIDES).numpy()
    np.testing.assert_allclose(pred, g["pred"], rtol=1e-6, atol=1e-6)


def _check_grads(g, p):
    for l, pi in enumerate(p):
        gr = pi.grad.numpy().reshape(-1)
        np.testing.assert_allclose(gr[g[f"g{l}_si"]], g[f"g{l}_sv"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(gr[g[f"g{l}_ti"]], g[f"g{l}_tv"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(np.abs(gr).sum(dtype=np.float64), float(g[f"g{l}_l1"]), rtol=1e-4)
        np.testing.assert_allclose(pi.grad.numpy()[..., 4].reshape(-1)[::7], g[f"g{l}_obj"], rtol=1e-4, atol=1e-8)


def test_sup_loss(golden):
    g = golden("loss_sup")
    B = int(g["B"])
    p = [torch.from_numpy(x).requires_grad_(True) for x in synth.make_head_logits(int(g["logit_seed"]), B)]
    tg = synth.make_targets(int(g["target_seed"]), int(g["n"]), B)
    sets = [port.build_targets(tg, synth.ANCHORS_GRID, synth.level_shapes())]
    loss, (lbox, lobj, lcls) = port.det_loss(p, sets, [4.0, 1.0, 0.4], 0.05, 0.7, 0.3)  # yaml Loss.box/obj/cls
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-5)
    np.testing.assert_allclose(lbox.detach().numpy(), g["box"], rtol=1e-5)
    np.testing.assert_allclose(lobj.detach().numpy(), g["obj"], rtol=1e-5)
    np.testing.assert_allclose(lcls.detach().numpy(), g["cls"], rtol=1e-5)
    loss.backward()
    _check_grads(g, p)


def test_ssod_loss(golden):
    g = golden("loss_ssod")
    B = int(g["B"])
    p = [torch.from_numpy(x).requires_grad_(True) for x in synth.make_head_logits(int(g["logit_seed"]), B)]
    rows = synth.make_pseudo_rows(int(g["rows_seed"]), int(g["n"]), B)
    sel = port.select_targets(rows, [0.6] * 80, [0.1] * 80, with_obj=True)
    shapes = synth.level_shapes()
    sets = [port.build_targets(sel[0][:, :6], synth.ANCHORS_GRID, shapes)] + \
           [port.build_targets(s, synth.ANCHORS_GRID, shapes, with_score=True) for s in sel[1:]]
    loss, (lbox, lobj, lcls) = port.det_loss(p, sets, [4.0, 1.0, 0.4], 0.05, 0.7, 0.3, with_bbox=True, with_cls=False)
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-5)
    np.testing.assert_allclose(lbox.detach().numpy(), g["box"], rtol=1e-5)
    np.testing.assert_allclose(lobj.detach().numpy(), g["obj"], rtol=1e-5)
    np.testing.assert_allclose(lcls.detach().numpy(), g["cls"], rtol=1e-5)
    loss.backward()
    _check_grads(g, p)


def test_ema(golden):
    g = golden("ema")
    keys = [k[len("init_"):] for k in g.files if k.startswith("init_")]
    import math
    ema = {k: g["init_" + k].copy() for k in keys}
    semi = {k: g["init_" + k].copy() for k in keys}
    ssup = {k: g["init_" + k].copy() for k in keys}
    for step in range(3):
        d = 0.9999 * (1 - math.exp(-(step + 1) / 2000))
        for k in keys:
            if ema[k].dtype.kind != "f":
                continue
            src = g[f"src{step}_{k}"]
            ema[k] = port.ema_update(ema[k], src, d)
            semi[k] = port.ema_update(semi[k], ema[k], float(g["semi_decay"]))
            ssup[k] = port.ema_update(ssup[k], src, 0.999)
            assert np.array_equal(ema[k], g[f"ema{step}_{k}"]), k       # bit-exact
            assert np.array_equal(semi[k], g[f"semi{step}_{k}"]), k
            assert np.array_equal(ssup[k], g[f"ssup{step}_{k}"]), k


@pytest.mark.parametrize("name", ["ml_cap", "ml_few", "best", "ml_agn"])
def test_val_path_nms_multi_label(golden, name):
    """non_max_suppression(multi_label=True, conf 0.001) including the 30 000-cap path.
    The oracle restatement is pinned here against the live reference; the CUDA kernel for it is the next widening step."""
    g = golden("nms_val")
    seed, B, P, frac, conf, iou, ml, agn = g[name + "_meta"]
    pred = synth.make_teacher_pred(int(seed), int(B), int(P), cand_frac=float(frac))
    if name == "ml_agn":
        pred[1, :, 4] = 0.0                # an image without candidates (class-agnostic case)
    got = port.nms_val(pred, float(conf), float(iou), multi_label=bool(ml), agnostic=bool(agn))
    for b in range(int(B)):
        want = g[f"{name}_det{b}"]
        assert got[b].shape == want.shape, (name, b, got[b].shape, want.shape)
        assert np.array_equal(got[b][:, 5], want[:, 5])                       # classes / order exact
        np.testing.assert_array_equal(got[b], want)                           # boxes and scores bit-exact


def test_labelmatch_bookkeeping_and_epoch_thresholds(golden):
    """the LabelMatch creator.  Its rows equal FairPseudoLabel's (asserted when the fixture was made), so
    the device pipeline is shared; here the host side of the mirror (score lists per class, per-epoch thresholds incl. the
    GMM split) is checked against the live reference, fed with the oracle's detections for the same seeded inputs."""
    from types import SimpleNamespace as NS
    from efficientteacher_b200.labelmatch import LabelMatch
    g = golden("labelmatch")
    assert bool(g["b0_same_as_fair"]) and bool(g["b1_same_as_fair"])
    cfg = NS(SSOD=NS(nms_conf_thres=float(g["nms_conf_thres"]), nms_iou_thres=float(g["nms_iou_thres"]), debug=False, multi_label=False,
                     ignore_thres_low=float(g["ignore_thres_low"]), ignore_thres_high=float(g["ignore_thres_high"]),
                     resample_high_percent=float(g["resample_high_percent"]), resample_low_percent=float(g["resample_low_percent"])),
             Dataset=NS(names=[str(i) for i in range(80)], np=0))
    lm = LabelMatch(cfg, 1000, 7.0, np.full(80, 1.0 / 80))
    for bi in range(2):
        seed, B, P, frac = g[f"b{bi}_meta"]
        pred = synth.make_teacher_pred(int(seed), int(B), int(P), cand_frac=float(frac))
        dets = port.nms_ssod(pred, lm.nms_conf_thres, lm.nms_iou_thres)
        det = np.zeros((int(B), 300, 8), np.float32)
        for b, d in enumerate(dets):
            det[b, :len(d)] = d
        lm.record_detections(det, np.array([len(d) for d in dets]))
        rows = port.pseudo_label_rows(dets, synth.make_Ms(int(seed) + 100, int(B)), 640, 640)
        np.testing.assert_allclose(rows, g[f"b{bi}_rows"], rtol=1e-9, atol=1e-9)
        lm.update(rows, n=int(B), pse_n=int(B))
    lens = np.array([len(c) for c in lm.score_list_epoch])
    assert np.array_equal(lens, g["epoch_score_lens"])
    assert np.array_equal(np.array([v for c in lm.score_list_epoch for v in c]), g["epoch_scores"])      # same values, same order
    assert np.array_equal(lm.cls_tmp, g["cls_tmp"])
    lm.update_epoch_cls_thr(0)
    np.testing.assert_allclose(np.array(lm.cls_thr_low), g["thr_low_e0"], rtol=0, atol=0)
    np.testing.assert_allclose(np.array(lm.cls_thr_high), g["thr_high_e0"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(lm.cls_num_total, g["cls_num_total_e0"]) and all(len(c) == 0 for c in lm.score_list_epoch)
