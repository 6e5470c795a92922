"""Generates tests/golden/*.npz from the LIVE, UNMODIFIED reference (imported from /root/reference through
oracle/ref_harness.py).  Run in the build container only:  python tests/golden/make_golden.py
Inputs are re-created from seeds by tests/synth.py, so only outputs (and small inputs) are stored."""
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

SSOD_YAML = 'configs/ssod/coco-standard/yolov5l_coco_ssod_10_percent.yaml'
SMALL = ['Model.depth_multiple', 0.33, 'Model.width_multiple', 0.50]   # YOLOv5s-sized model: same head/anchors


def pack_assign(res, with_score):
    if with_score:
        tcls, tbox, indices, anch, tscore = res
    else:
        tcls, tbox, indices, anch = res
        tscore = [torch.zeros(len(c)) for c in tcls]
    d = {}
    for l in range(len(tcls)):
        d[f"idx{l}"] = torch.stack(indices[l], 1).numpy().astype(np.int64)
        d[f"tbox{l}"] = tbox[l].numpy()
        d[f"anch{l}"] = anch[l].numpy()
        d[f"tcls{l}"] = tcls[l].numpy().astype(np.int64)
        d[f"tscore{l}"] = tscore[l].numpy()
    return d


def sample_idx(n, k, seed):
    return np.random.RandomState(seed).randint(0, n, k)


def gen_val_nms(ns):
    """val.py-path NMS (general.py:994-1098): multi_label at conf 0.001 (the 30000-cap path) and at 0.25, best-class twin."""
    cases = (("ml_cap", 41, 2, 2000, 0.3, 0.001, 0.6, True, False), ("ml_few", 42, 3, 4000, 0.05, 0.25, 0.45, True, False),
             ("best", 43, 2, 4000, 0.05, 0.001, 0.6, False, False), ("ml_agn", 44, 3, 1500, 0.1, 0.05, 0.5, True, True))
    out = {}
    for name, seed, B, P, frac, conf, iou, ml, agn in cases:
        pred = synth.make_teacher_pred(seed, B, P, cand_frac=frac)
        if name == "ml_agn":
            pred[1, :, 4] = 0.0            # an image without candidates
        dets = ns.non_max_suppression(torch.from_numpy(pred).clone(), conf_thres=conf, iou_thres=iou, multi_label=ml, agnostic=agn)
        out[name + "_meta"] = np.array([seed, B, P, frac, conf, iou, float(ml), float(agn)], dtype=np.float64)
        for b in range(B):
            out[f"{name}_det{b}"] = dets[b].numpy().reshape(-1, 6)
        print("val nms", name, [len(d) for d in dets])
    np.savez_compressed(os.path.join(HERE, "nms_val.npz"), **out)


def gen_labelmatch(ns):
    """utils/labelmatch.py: the LabelMatch pseudo-label creator over two batches + the per-epoch threshold update (GMM)."""
    import importlib
    lm = importlib.import_module("utils.labelmatch")
    cfg = ref_harness.make_cfg(SSOD_YAML, SMALL)
    nc = 80
    torch.manual_seed(0)
    np.random.seed(0)
    creator = lm.LabelMatch(cfg, 1000, 7.0, np.full(nc, 1.0 / nc))
    fpl = ns.FairPseudoLabel(cfg)
    out = dict(nms_conf_thres=creator.nms_conf_thres, nms_iou_thres=creator.nms_iou_thres,
               resample_high_percent=cfg.SSOD.resample_high_percent, resample_low_percent=cfg.SSOD.resample_low_percent,
               ignore_thres_high=cfg.SSOD.ignore_thres_high, ignore_thres_low=cfg.SSOD.ignore_thres_low)
    for bi, (seed, B, P, frac) in enumerate(((51, 4, 25200, 0.02), (52, 4, 25200, 0.03))):
        pred = synth.make_teacher_pred(seed, B, P, cand_frac=frac)
        Ms = synth.make_Ms(seed + 100, B)
        imgs = torch.zeros(B, 3, 640, 640)
        rows, invalid = creator.create_pseudo_label_online_with_gt(torch.from_numpy(pred).clone(), imgs, torch.from_numpy(Ms), imgs.clone())
        rows = rows.numpy() if isinstance(rows, torch.Tensor) else np.zeros((0, 9))
        rows_f, _ = fpl.create_pseudo_label_online_with_gt(torch.from_numpy(pred).clone(), imgs, torch.from_numpy(Ms), imgs.clone())
        rows_f = rows_f.numpy() if isinstance(rows_f, torch.Tensor) else np.zeros((0, 9))
        print("labelmatch batch", bi, rows.shape, "identical to FairPseudoLabel rows:", rows.shape == rows_f.shape and np.array_equal(rows, rows_f))
        out[f"b{bi}_meta"] = np.array([seed, B, P, frac], dtype=np.float64)
        out[f"b{bi}_rows"] = rows
        out[f"b{bi}_same_as_fair"] = np.array(rows.shape == rows_f.shape and np.array_equal(rows, rows_f))
        creator.update(rows, n=B, pse_n=B)
    lens = np.array([len(c) for c in creator.score_list_epoch], dtype=np.int64)
    out["epoch_score_lens"] = lens
    out["epoch_scores"] = np.array([v for c in creator.score_list_epoch for v in c], dtype=np.float64)
    out["cls_tmp"] = creator.cls_tmp.copy()
    creator.update_epoch_cls_thr(0)
    out["thr_high_e0"] = np.array(creator.cls_thr_high, dtype=np.float64)
    out["thr_low_e0"] = np.array(creator.cls_thr_low, dtype=np.float64)
    out["cls_num_total_e0"] = creator.cls_num_total.copy()
    np.savez_compressed(os.path.join(HERE, "labelmatch.npz"), **out)
    print("labelmatch thr_high[:8]", out["thr_high_e0"][:8], "thr_low[:8]", out["thr_low_e0"][:8])


def gen_val_process_batch(ns):
    """val.py:123-145 process_batch of the live reference on seeded detections / labels (no exact IoU ties)"""
    os.environ.setdefault("WANDB_MODE", "disabled")
    import val as V
    r = np.random.RandomState(11)
    cases = {}
    for name, (N, M) in {"a": (300, 40), "b": (50, 120), "c": (7, 1), "d": (200, 200)}.items():
        lab = np.zeros((M, 5), np.float32)
        lab[:, 0] = r.randint(0, 5, M)
        xy = r.uniform(50, 550, (M, 2)); wh = r.uniform(20, 200, (M, 2))
        lab[:, 1:3] = xy - wh / 2; lab[:, 3:5] = xy + wh / 2
        det = np.zeros((N, 6), np.float32)
        src = r.randint(0, M, N)
        det[:, :4] = lab[src, 1:5] + r.normal(0, 12, (N, 4)).astype(np.float32)
        det[:, 4] = np.sort(r.uniform(0.001, 1, N))[::-1]
        det[:, 5] = np.where(r.uniform(size=N) < 0.8, lab[src, 0], r.randint(0, 5, N))
        want = V.process_batch(torch.from_numpy(det), torch.from_numpy(lab), torch.linspace(0.5, 0.95, 10)).numpy()
        cases[name + "_det"] = det; cases[name + "_lab"] = lab; cases[name + "_correct"] = want
    np.savez_compressed(os.path.join(HERE, "val_process_batch.npz"), **cases)


def gen_extra_teachers(ns):
    """FairPseudoLabel.create_pseudo_label_online_with_extra_teachers (self_supervised_utils.py:249-313) up to the point where
    the reference raises: output_to_target_ssod unpacks 8 columns from the 6-column rows (plots.py:488).  The harness
    intercepts that call and keeps its argument = the merged per-image detections."""
    import utils.self_supervised_utils as U
    fp = ns.FairPseudoLabel(ref_harness.make_cfg(SSOD_YAML))
    captured = {}

    class _Stop(Exception):
        pass

    def capture(out):
        captured["out"] = [o.clone() for o in out]
        raise _Stop()
    orig, U.output_to_target_ssod = U.output_to_target_ssod, capture
    try:
        B, P = 2, 4000
        pred = synth.make_teacher_pred(21, B, P, cand_frac=0.05)
        e1 = synth.make_teacher_pred(22, B, P, cand_frac=0.04)
        e2 = synth.make_teacher_pred(23, B, P, cand_frac=0.03)
        try:
            fp.create_pseudo_label_online_with_extra_teachers(torch.from_numpy(pred), [torch.from_numpy(e1), torch.from_numpy(e2)],
                                                              torch.zeros(B, 3, 640, 640), torch.from_numpy(synth.make_Ms(9, B, 640)),
                                                              [{3: 70, 5: 1, 7: 7}, {}], -1)
        except _Stop:
            pass
    finally:
        U.output_to_target_ssod = orig
    ref = [o.numpy() for o in captured["out"]]
    np.savez_compressed(os.path.join(HERE, "extra_teachers.npz"), conf=np.float32(fp.nms_conf_thres), iou=np.float32(fp.nms_iou_thres),
                        n0=len(ref[0]), out0=ref[0], out1=ref[1])


def main():
    ns = ref_harness.load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "--only-val-nms":      # adds one fixture without touching the others
        torch.set_num_threads(8)
        gen_val_nms(ns)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--only-val-extra":      # val.process_batch + extra-teachers merge fixtures
        gen_val_process_batch(ns)
        gen_extra_teachers(ns)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--only-labelmatch":
        torch.set_num_threads(8)
        gen_labelmatch(ns)
        return
    torch.set_num_threads(8)
    cfg = ref_harness.make_cfg(SSOD_YAML, SMALL)
    torch.manual_seed(0)
    model = ns.SSODModel(cfg)
    det = model.head
    assert np.allclose(det.anchors.numpy(), synth.ANCHORS_GRID)

    # ---- build_targets / build_uc_targets_aug ----
    assigner = ns.YOLOAnchorAssigner(det.na, det.nl, det.anchors, cfg.Loss.anchor_t, det.stride, det.nc, 0)
    for name, B, n in (("a", 16, 128), ("b", 2, 3), ("c", 4, 1000)):
        p = [torch.zeros(B, 3, ny, nx, 85) for ny, nx in synth.level_shapes()]
        t = synth.make_targets(11, n, B)
        d = pack_assign(assigner(p, torch.from_numpy(t)), False)
        sc = np.random.RandomState(5).uniform(0.1, 1, (n, 1)).astype(np.float32)
        t7 = np.concatenate([t, sc], 1)
        d7 = pack_assign(assigner(p, torch.from_numpy(t7), with_pseudo_score=True), True)
        np.savez_compressed(os.path.join(HERE, f"assign_{name}.npz"), B=B, n=n, seed=11, score_seed=5,
                            **{"bt_" + k: v for k, v in d.items()}, **{"uc_" + k: v for k, v in d7.items()})
    # empty targets
    p = [torch.zeros(2, 3, ny, nx, 85) for ny, nx in synth.level_shapes()]
    d = pack_assign(assigner(p, torch.zeros(0, 6)), False)
    assert all(len(d[f"tcls{l}"]) == 0 for l in range(3))

    # ---- NMS + pseudo labels ----
    fpl = ns.FairPseudoLabel(cfg)
    cases = (("a", 21, 4, 25200, 0.02), ("dense", 22, 2, 3000, 0.5), ("cap", 23, 1, 42000, 0.95), ("hi", 24, 2, 100800, 0.02))
    for name, seed, B, P, frac in cases:
        pred = synth.make_teacher_pred(seed, B, P, cand_frac=frac)
        if name == "a":
            pred[2, :, 4] = 0.01          # an image without candidates
        Ms = synth.make_Ms(seed + 100, B)
        tp = torch.from_numpy(pred)
        dets = ns.non_max_suppression_ssod(tp.clone(), conf_thres=fpl.nms_conf_thres, iou_thres=fpl.nms_iou_thres)
        imgs = torch.zeros(B, 3, 640, 640)
        rows, invalid = fpl.create_pseudo_label_online_with_gt(tp.clone(), imgs, torch.from_numpy(Ms), imgs.clone())
        rows = rows.numpy() if isinstance(rows, torch.Tensor) else np.zeros((0, 9))
        out = dict(seed=seed, B=B, P=P, frac=frac, conf_thres=fpl.nms_conf_thres, iou_thres=fpl.nms_iou_thres,
                   rows=rows, invalid=invalid, Ms=Ms)
        for b in range(B):
            out[f"det{b}"] = dets[b].numpy().reshape(-1, 8)
        if name in ("a", "dense"):
            d6 = ns.non_max_suppression(tp.clone(), conf_thres=0.25, iou_thres=0.45)
            for b in range(B):
                out[f"val{b}"] = d6[b].numpy().reshape(-1, 6)
        np.savez_compressed(os.path.join(HERE, f"nms_{name}.npz"), **out)
        print("nms", name, [len(d) for d in dets], rows.shape)

    # ---- select_targets ----
    ssod_loss = ns.ComputeStudentMatchLoss(model, cfg)
    rows = synth.make_pseudo_rows(31, 600, 8)
    ssod_loss.ignore_thres_high[3] = 0.8
    ssod_loss.ignore_thres_low[5] = 0.3
    sel = ssod_loss.select_targets(torch.from_numpy(rows))
    np.savez_compressed(os.path.join(HERE, "select.npz"), seed=31, n=600, B=8, high=np.array(ssod_loss.ignore_thres_high),
                        low=np.array(ssod_loss.ignore_thres_low), **{f"s{i}": s.numpy().reshape(-1, 7) for i, s in enumerate(sel)})
    ssod_loss.ignore_thres_high[3] = 0.6
    ssod_loss.ignore_thres_low[5] = 0.1

    # ---- losses (forward values + gradient samples) ----
    sup_loss = ns.ComputeLoss(model, cfg)
    B = 4
    logits = synth.make_head_logits(41, B)
    tg = synth.make_targets(42, 64, B)
    p = [torch.from_numpy(x).requires_grad_(True) for x in logits]
    loss, items = sup_loss(p, torch.from_numpy(tg))
    loss.backward()
    out = dict(B=B, logit_seed=41, target_seed=42, n=64, loss=loss.detach().numpy(),
               box=items["box"].detach().numpy(), obj=items["obj"].detach().numpy(), cls=items["cls"].detach().numpy())
    for l, pi in enumerate(p):
        g = pi.grad.numpy().reshape(-1)
        si = sample_idx(len(g), 4096, 50 + l)
        top = np.argsort(-np.abs(g))[:512]
        out[f"g{l}_l1"] = np.abs(g).sum(dtype=np.float64)
        out[f"g{l}_si"], out[f"g{l}_sv"] = si, g[si]
        out[f"g{l}_ti"], out[f"g{l}_tv"] = top, g[top]
        out[f"g{l}_obj"] = pi.grad.numpy()[..., 4].reshape(-1)[::7].copy()
    np.savez_compressed(os.path.join(HERE, "loss_sup.npz"), **out)
    print("sup loss", float(loss), {k: float(v) for k, v in items.items()})

    rows = synth.make_pseudo_rows(43, 400, B)
    p = [torch.from_numpy(x).requires_grad_(True) for x in logits]
    loss, items = ssod_loss(p, torch.from_numpy(rows))
    loss.backward()
    out = dict(B=B, logit_seed=41, rows_seed=43, n=400, loss=loss.detach().numpy(), box=items["ss_box"].detach().numpy(),
               obj=items["ss_obj"].detach().numpy(), cls=items["ss_cls"].detach().numpy())
    for l, pi in enumerate(p):
        g = pi.grad.numpy().reshape(-1)
        si = sample_idx(len(g), 4096, 60 + l)
        top = np.argsort(-np.abs(g))[:512]
        out[f"g{l}_l1"] = np.abs(g).sum(dtype=np.float64)
        out[f"g{l}_si"], out[f"g{l}_sv"] = si, g[si]
        out[f"g{l}_ti"], out[f"g{l}_tv"] = top, g[top]
        out[f"g{l}_obj"] = pi.grad.numpy()[..., 4].reshape(-1)[::7].copy()
    np.savez_compressed(os.path.join(HERE, "loss_ssod.npz"), **out)
    print("ssod loss", float(loss), {k: float(v) for k, v in items.items()})

    # ---- bbox_iou CIoU ----
    r = np.random.RandomState(71)
    b1 = np.concatenate([r.uniform(-0.5, 1.5, (512, 2)), r.uniform(0.05, 8, (512, 2))], 1).astype(np.float32)
    b2 = np.concatenate([r.uniform(0, 1, (512, 2)), r.uniform(0.05, 8, (512, 2))], 1).astype(np.float32)
    c = ns.bbox_iou(torch.from_numpy(b1).T, torch.from_numpy(b2), x1y1x2y2=False, CIoU=True).numpy()
    np.savez_compressed(os.path.join(HERE, "ciou.npz"), b1=b1, b2=b2, ciou=c)

    # ---- Detect eval decode ----
    det.eval()
    r = np.random.RandomState(81)
    feats = [torch.from_numpy(r.standard_normal((2, c, s, s)).astype(np.float32)) for c, s in zip((128, 256, 512), (8, 4, 2))]
    with torch.no_grad():
        raw = [m(f) for m, f in zip(det.m, feats)]
        raw5 = [x.view(2, 3, 85, x.shape[2], x.shape[3]).permute(0, 1, 3, 4, 2).contiguous() for x in raw]
        z, _ = det(list(feats))
    np.savez_compressed(os.path.join(HERE, "decode.npz"), pred=z.numpy(), **{f"raw{l}": x.numpy() for l, x in enumerate(raw5)})

    # ---- EMA (3 classes, a few updates, odd tensor sizes incl. an int buffer) ----
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Conv2d(3, 7, 3, bias=False)
            self.bn = torch.nn.BatchNorm2d(7)
            self.b = torch.nn.Linear(13, 5)
    torch.manual_seed(5)
    src = Tiny()
    ema = ns.ModelEMA(src)
    semi = ns.CosineEMA(ema.ema, decay_start=0.99, decay_end=0.9999, total_epoch=10)
    semi.update_decay(3)
    ssup = ns.SemiSupModelEMA(src, 0.999)
    out = {}
    out.update({f"init_{k}": v.numpy().copy() for k, v in src.state_dict().items()})
    for step in range(3):
        with torch.no_grad():
            g = torch.Generator().manual_seed(100 + step)
            for k, v in src.state_dict().items():
                if v.dtype.is_floating_point:
                    v += torch.randn(v.shape, generator=g) * 0.1
        out.update({f"src{step}_{k}": v.numpy().copy() for k, v in src.state_dict().items()})
        ema.update(src); semi.update(ema.ema); ssup.update(src)
        out.update({f"ema{step}_{k}": v.numpy().copy() for k, v in ema.ema.state_dict().items()})
        out.update({f"semi{step}_{k}": v.numpy().copy() for k, v in semi.ema.state_dict().items()})
        out.update({f"ssup{step}_{k}": v.numpy().copy() for k, v in ssup.ema.state_dict().items()})
    out["semi_decay"] = semi.decay
    np.savez_compressed(os.path.join(HERE, "ema.npz"), **out)
    gen_val_process_batch(ns)
    gen_extra_teachers(ns)
    print("done")


if __name__ == "__main__":
    main()
