"""Why do the EMA teacher's logits move by tens of percent after ONE optimizer step at decay 0.9999?  (tools/debug_teacher_drift.py
prints that drift; the engine itself is deterministic.)  Separates: (a) how much each teacher tensor changed, (b) whether
restoring the parameters restores the logits, (c) what an independent fp32 torch forward (oracle/trunk_ref.py on the GPU) says
about the sensitivity of the logits to that parameter change."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import __graft_entry__ as g
    g.build()
    import synth
    from efficientteacher_b200.config import yolov5_ssod_cfg
    from efficientteacher_b200.trainer import SSODTrainerStep
    from oracle.trunk_ref import TrunkRef
    dev = torch.device("cuda", 0)
    size, img, bl, bu = "l", 640, 8, 8
    torch.manual_seed(0)
    cfg = yolov5_ssod_cfg(size, batch_size=bl + bu, img_size=img)
    cfg.SSOD.fixed_accumulate = True
    st = SSODTrainerStep(cfg, dev, epochs=300)
    st.ema.updates = 100000
    r = np.random.RandomState(3)
    imgs = torch.from_numpy(r.rand(bl, 3, img, img).astype(np.float32)).to(dev)
    uw = torch.from_numpy(r.rand(bu, 3, img, img).astype(np.float32)).to(dev)
    us = uw.flip(3).contiguous()
    tg = torch.from_numpy(synth.make_targets(7, 8 * bl, bl)).to(dev)
    Ms = torch.from_numpy(synth.make_Ms(9, bu, img)).to(dev)
    with torch.no_grad():
        bns = [m for m in st.model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
        for m in bns:
            m.momentum = 1.0
        with torch.autocast("cuda", dtype=torch.bfloat16):
            st.model(torch.cat([imgs, us], 0).contiguous(memory_format=torch.channels_last))
        for m in bns:
            m.momentum = 0.03
        for h in st.model.head.m:
            h.bias.view(3, -1)[:, 4] += 6.5
            h.bias.view(3, -1)[:, 5:] += 5.0
        st.ema.ema.load_state_dict(st.model.state_dict())
        st.semi_ema.ema.load_state_dict(st.model.state_dict())
        sd0 = {k: v.detach().clone() for k, v in st.ema.ema.state_dict().items()}
        (_, raw0), _ = st.ema.ema(uw)
        raw0 = [t.clone() for t in raw0]
        ref0, _ = TrunkRef(sd0, (3, 6, 9, 3), 3).forward(uw, train=False)
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))  # noqa: E731
    print("engine vs fp32 torch on the initial teacher, per level:", ["%.4g" % rel(a, b) for a, b in zip(raw0, ref0)], flush=True)
    for step in range(2):
        st.train_instance(imgs, tg, us, uw, None, Ms, step)
        with torch.no_grad():
            sd1 = {k: v.detach().clone() for k, v in st.ema.ema.state_dict().items()}
            (_, raw1), _ = st.ema.ema(uw)
            raw1 = [t.clone() for t in raw1]
            ref1, _ = TrunkRef(sd1, (3, 6, 9, 3), 3).forward(uw, train=False)
            rows = sorted(((rel(sd1[k].float(), sd0[k].float()), float((sd1[k].float() - sd0[k].float()).abs().max()), k) for k in sd0 if sd0[k].dtype.is_floating_point), reverse=True)
            print("after step %d: engine drift %s | fp32-torch drift %s" % (step, ["%.4g" % rel(a, b) for a, b in zip(raw1, raw0)],
                                                                            ["%.4g" % rel(a, b) for a, b in zip(ref1, ref0)]), flush=True)
            print("   largest relative teacher-tensor changes:", ["%s %.3g (abs %.3g)" % (k, a, b) for a, b, k in rows[:6]], flush=True)
            # put the initial parameters back: do the logits come back?
            st.ema.ema.load_state_dict(sd0)
            (_, raw2), _ = st.ema.ema(uw)
            print("   parameters restored -> engine drift %s" % ["%.4g" % rel(a, b) for a, b in zip(raw2, raw0)], flush=True)
            st.ema.ema.load_state_dict(sd1)
            # which group of tensors carries the sensitivity?  apply the change of one group at a time (fp32 torch forward)
            for name, pred in (("bn.weight", lambda k: k.endswith("bn.weight")), ("running_mean", lambda k: k.endswith("running_mean")),
                               ("running_var", lambda k: k.endswith("running_var")), ("conv/head weights+biases", lambda k: not ("bn." in k))):
                mix = {k: (sd1[k] if pred(k) else sd0[k]) for k in sd0}
                refm, _ = TrunkRef(mix, (3, 6, 9, 3), 3).forward(uw, train=False)
                print("   only %-26s changed -> fp32-torch drift %s" % (name, ["%.4g" % rel(a, b) for a, b in zip(refm, ref0)]), flush=True)


if __name__ == "__main__":
    main()
