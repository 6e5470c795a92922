"""Bisect tool for the open issue in DESIGN.md section 7 item 0 (teacher collapse in the synthetic bench).

Runs K SSOD steps of the bench workload (or a smaller one) and prints, per step: loss, relative drift of the teacher's raw
logits on the fixed unlabeled batch, BN extrema of student and teacher, and the largest |grad| of every optimizer group.
Switches turn individual fast paths off so one GPU call can compare variants:

    python tools/debug_teacher_drift.py --size l --img 640 --bl 16 --bu 16 --steps 16 --variants base,noside,noglue,nofanin,eager

variants: base (shipping defaults, graphed) | eager (no CUDA graph) | noside (ETB_WGRAD_SIDE=0) | nofanin | noglue |
          torchbn (Conv.FUSED_BN=False: torch BatchNorm + SiLU around the native convs) | torchconv (Conv.NATIVE=False)
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def bn_ext(mod):
    bb = [q for q in mod.modules() if isinstance(q, torch.nn.BatchNorm2d)]
    return (max(float(q.running_var.max()) for q in bb), max(float(q.running_mean.abs().max()) for q in bb),
            max(float(q.weight.abs().max()) for q in bb), max(float(q.bias.abs().max()) for q in bb))


def bn_top(mod, k=3):
    rows = [(float(q.running_var.max()), float(q.weight.abs().max()), n) for n, q in mod.named_modules() if isinstance(q, torch.nn.BatchNorm2d)]
    rows.sort(reverse=True)
    return " ".join("%s(var %.3g,|g| %.3g)" % (n, v, g) for v, g, n in rows[:k])


def run(variant, args):
    import synth
    from efficientteacher_b200 import autograd_conv as ac
    from efficientteacher_b200.config import yolov5_ssod_cfg
    from efficientteacher_b200.model import Conv
    from efficientteacher_b200.trainer import SSODTrainerStep
    Conv.NATIVE, Conv.FUSED_BN, Conv.FUSED_GLUE, Conv.FUSED_FANIN = True, True, True, True
    SSODTrainerStep.WGRAD_SIDE_STREAM = True
    graph = True
    for v in variant.split("+"):
        if v == "eager":
            graph = False
        elif v == "noside":
            SSODTrainerStep.WGRAD_SIDE_STREAM = False
        elif v == "nofanin":
            Conv.FUSED_FANIN = False
        elif v == "noglue":
            Conv.FUSED_GLUE = False
        elif v == "torchbn":
            Conv.FUSED_BN = False
            graph = False
        elif v == "torchconv":
            Conv.NATIVE = False
            graph = False
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    cfg = yolov5_ssod_cfg(args.size, batch_size=args.bl + args.bu, img_size=args.img)
    cfg.SSOD.fixed_accumulate = True
    st = SSODTrainerStep(cfg, dev, epochs=300)
    if args.nw is not None:
        st.nw = args.nw           # 0 = the round-1 regime (full lr on a random init), default = the reference's warm-up (1000)
    st.ema.updates = args.updates
    r = np.random.RandomState(3)
    imgs = torch.from_numpy(r.rand(args.bl, 3, args.img, args.img).astype(np.float32)).to(dev)
    uw = torch.from_numpy(r.rand(args.bu, 3, args.img, args.img).astype(np.float32)).to(dev)
    us = uw.flip(3).contiguous()
    tg = torch.from_numpy(synth.make_targets(7, 8 * args.bl, args.bl)).to(dev)
    Ms = torch.from_numpy(synth.make_Ms(9, args.bu, args.img)).to(dev)
    with torch.no_grad():
        bns = [m for m in st.model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
        for m in bns:
            m.momentum = 1.0
        with torch.autocast("cuda", dtype=torch.bfloat16):
            st.model(torch.cat([imgs, us], 0).contiguous(memory_format=torch.channels_last))
        for m in bns:
            m.momentum = 0.03
        for h in st.model.head.m:           # candidates for the pseudo-label path (objectness ~0.5, class scores ~0.5)
            h.bias.view(3, -1)[:, 4] += 6.5
            h.bias.view(3, -1)[:, 5:] += 5.0
        st.ema.ema.load_state_dict(st.model.state_dict())
        st.semi_ema.ema.load_state_dict(st.model.state_dict())
        (_, raw0), _ = st.ema.ema(uw)
        raw0 = [t.clone() for t in raw0]
        # determinism self-check of the teacher engine: the same forward twice more, per level
        for rep in range(2):
            (_, rawr), _ = st.ema.ema(uw)
            print("   teacher forward repeat %d: per-level rel diff vs first call %s, |raw0| %s" % (
                rep, ["%.3g" % float((a - b).norm() / b.norm()) for a, b in zip(rawr, raw0)], ["%.4g" % float(b.norm()) for b in raw0]), flush=True)
    groups = [[p for p in g["params"]] for g in st.optimizer.param_groups]
    print("== variant %s (graph=%s) updates=%d" % (variant, graph, args.updates), flush=True)
    for i in range(args.steps):
        if graph:
            loss = st.train_instance_graphed(imgs, tg, us, uw, None, Ms, i)
            gmax = ["-"] * 3                      # the arena is zeroed inside the captured optimizer step
        else:
            loss = st.train_instance(imgs, tg, us, uw, None, Ms, i, _stop_after_backward=True)
            gmax = ["%.3g" % max(float(p.grad.abs().max()) for p in g if p.grad is not None) for g in groups]
            st._allreduce_grads()
            st._optimizer_ema(i)
        with torch.no_grad():
            (_, raw1), _ = st.ema.ema(uw)
        drifts = [float((a - b).norm() / b.norm()) for a, b in zip(raw1, raw0)]
        drift = max(drifts)
        if i < 2:
            print("        per-level drift %s" % ["%.4g" % d for d in drifts], flush=True)
        print("step %2d loss %.4f rows %d teacher drift %.5f | student BN (var,|mean|,|g|,|b|) %s | teacher %s | max|grad| bias/w/bnw %s" % (
            i, float(loss), int(st.pseudo_label_creator.last_count_dev.item()), drift, "%.3g %.3g %.3g %.3g" % bn_ext(st.model), "%.3g %.3g %.3g %.3g" % bn_ext(st.ema.ema), gmax), flush=True)
        if i in (0, args.steps - 1) or drift > 0.05:
            wmax = max((float(p.abs().max()), n) for n, p in st.model.named_parameters() if p.dim() == 4)
            print("        student top BN: %s | largest conv weight %.3g (%s)" % (bn_top(st.model), wmax[0], wmax[1]), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="l_shallow")
    ap.add_argument("--img", type=int, default=256)
    ap.add_argument("--bl", type=int, default=4)
    ap.add_argument("--bu", type=int, default=4)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--updates", type=int, default=100000)
    ap.add_argument("--nw", type=int, default=None)
    ap.add_argument("--variants", default="base,eager,noside,nofanin,noglue,torchbn")
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    for v in args.variants.split(","):
        run(v, args)


if __name__ == "__main__":
    main()
