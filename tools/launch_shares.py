"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_shares.py list.csv [iterations]
With `iterations` the time per iteration is printed too (the small-workload lists hold a whole number of iterations)."""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = OrderedDict()
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"^void |\(.*$|htlr::", "", r[ki])
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    t = acc.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += v
tot = sum(v for _, v in acc.values())
n = sum(c for c, _ in acc.values())
its = int(sys.argv[2]) if len(sys.argv) > 2 else 0
print(f"launches {n}, total kernel time {tot / 1e3:.2f} ms" + (f" = {its} iterations, {tot / its:.1f} us of kernel time per iteration" if its else ""))
for name, (c, v) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    per = f"  {v / its:8.1f} us per iteration" if its else f"  {v:12.1f} us"
    print(f"{100 * v / tot:6.2f} %  {c:5d} launches{per}  {name}")
