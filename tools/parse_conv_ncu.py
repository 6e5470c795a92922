import csv, sys, json
sys.path.insert(0, 'tools')
from conv_bench import SHAPES
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
def load(m):
    lines = open('gpurun_out/conv_launches_%s.csv' % m).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    return list(csv.DictReader(lines[start:]))
R = {m: load(m) for m in ('fwd', 'dgrad', 'wgrad')}
idx = {m: 0 for m in R}
tot = {m: [0.0, 0.0] for m in R}
out = []
for name, Cin, Cout, k, s, H, cnt in SHAPES:
    p = k // 2; Ho = (H + 2 * p - k) // s + 1
    fl = 2.0 * N * Ho * Ho * Cout * Cin * k * k
    line = "%-20s x%-2d %6.1fGF" % (name, cnt, fl / 1e9)
    for m in ('fwd', 'dgrad', 'wgrad'):
        if m == 'dgrad' and Cout % 64: line += " | dgrad   -   "; continue
        nl = 4 if (m == 'dgrad' and s == 2) else 1
        rows = R[m][idx[m]:idx[m] + 3 * nl]; idx[m] += 3 * nl
        t = sum(float(x['Metric Value'].replace(',', '')) for x in rows[-nl:]) / 1e3
        tot[m][0] += t * cnt; tot[m][1] += fl * cnt
        line += " | %s %6.1fus %4.0fTF g%s" % (m, t, fl / t / 1e6, rows[-1]['Grid Size'].replace(' ', ''))
    print(line)
print({m: (idx[m], len(R[m])) for m in R})
for m, (t, f) in tot.items():
    print(m, "trunk total %.2f ms  %.0f TF/s" % (t / 1e3, f / t / 1e6))
