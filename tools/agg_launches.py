#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: usage agg_launches.py FILE [last|all]
`last` (default) = the launches from the last staging (prep_count_kernel) on, i.e. one reduction of one step."""
import collections, csv, re, sys
path = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "last"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
if mode == "last":
    idx = [i for i, x in enumerate(rows) if "prep_count" in x["Kernel Name"]]
    rows = rows[idx[-1]:]
agg, tot = collections.OrderedDict(), 0.0
for x in rows:
    n = x["Kernel Name"]
    m = re.search(r"limb_gemm_tc_kernel<(\d+), *(\d+)>", n)
    n = f"limb_gemm_tc_kernel<{m.group(1)},{['RMW','RMW_LIMBS','LIMBS'][int(m.group(2))]}>" if m else re.sub(r"\(.*", "", n)[:70]
    v = float(x["Metric Value"]) / 1e3
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += v; tot += v
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:70s} {c:5d} {v:10.1f} us {100 * v / tot:5.1f}%")
print(f"total {tot:.1f} us over {len(rows)} launches")
