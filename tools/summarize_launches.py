"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel markdown table (launches, total us, share)."""
import csv
import gzip
import re
import sys


def main(path, out, title):
    op = gzip.open if path.endswith(".gz") else open
    rows = []
    with op(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        rows.append((r[ki], us))
    agg = {}
    for k, us in rows:
        k = re.sub(r"\(.*", "", k)[:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    ours = ("conv_", "wgrad", "bn_", "stem_", "pack_", "fold_", "maxpool5", "upsample2x", "copy_slice", "sppf", "detect_", "loss_", "ema_",
            "sgd_", "cand_", "rank_", "nms_", "pl_", "build_", "select_", "nchw", "nhwc", "ciou")
    mine = sum(v[1] for k, v in agg.items() if any(s in k for s in ours))
    with open(out, "w") as f:
        f.write("# %s\n\nkernels in the step: %d, sum of durations: %.2f ms (cold-cache, serialised under ncu: compare SHARES)\n\n" % (title, len(rows), tot / 1e3))
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, n, us, 100 * us / tot))
        f.write("\nhand-written kernels (libetb200.so): %.1f%% of the step's GPU time; ATen / cuDNN / NCCL: %.1f%%\n" % (100 * mine / tot, 100 - 100 * mine / tot))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu launch list")
