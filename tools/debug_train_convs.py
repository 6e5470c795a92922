"""Diagnostic: run one student forward/backward and, inside every native conv backward, compare dx / dw with torch's
conv2d_input / conv2d_weight on the same (bf16-rounded) tensors.  Prints the layers that disagree."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import __graft_entry__ as g
    g.build()
    import synth
    from efficientteacher_b200 import autograd_conv as AC, model as M
    from efficientteacher_b200.config import yolov5_ssod_cfg
    from efficientteacher_b200.loss import ComputeLoss
    dev = "cuda:0"
    torch.manual_seed(0)
    cfg = yolov5_ssod_cfg('l_shallow', batch_size=4, img_size=256)
    m = M.Model(cfg).to(dev).train()
    crit = ComputeLoss(m, cfg)
    orig_bwd = AC.ConvFn.backward
    report = []

    def checked(ctx, dy):
        dx, dw, _, _ = orig_bwd(ctx, dy)
        x, w = ctx.saved_tensors
        s, p = ctx.geom
        bf = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
        msg = "w%s x%s xs%s dy%s dys%s s%d" % (tuple(w.shape), tuple(x.shape), x.stride(), tuple(dy.shape), dy.stride(), s)
        if dx is not None:
            ref = torch.nn.grad.conv2d_input(x.shape, bf(w), bf(dy), stride=s, padding=p)
            e = (dx.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
            report.append(("dx", e, msg))
        if dw is not None:
            ref = torch.nn.grad.conv2d_weight(bf(x), w.shape, bf(dy), stride=s, padding=p)
            e = (dw - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
            report.append(("dw", e, msg))
        return dx, dw, None, None

    AC.ConvFn.backward = staticmethod(checked)
    x = torch.rand(4, 3, 256, 256, generator=torch.Generator().manual_seed(1)).to(dev)
    tg = torch.from_numpy(synth.make_targets(2, 32, 4)).to(dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred, feat = m(x.contiguous(memory_format=torch.channels_last))
    loss, _ = crit([p.float() for p in pred], tg)
    (loss + sum(f.float().mean() for f in feat) * 0.1).backward()
    torch.cuda.synchronize()
    bad = [r for r in report if r[1] > 0.03]
    print("checked", len(report), "bad", len(bad))
    for kind, e, msg in bad[:40]:
        print("%s rel_err=%.3f %s" % (kind, e, msg))


if __name__ == "__main__":
    main()
