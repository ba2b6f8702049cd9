"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of ONE local step
(delimited by two consecutive gather_normalize launches).  Usage: summarize_launches.py launches.csv [title]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    try:
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r.get("Grid Size", "")))
    except Exception:  # noqa: BLE001
        pass
idx = [i for i, r in enumerate(rows) if "gather_normalize" in r[0]]
seg = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, _ in seg:
    key = re.sub(r"\(.*", "", re.sub(r"<.*", "", n)).replace("void ", "")[:80]
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v for _, v in agg.values())
ours = sum(v for k, (_, v) in agg.items() if k.startswith("rlr::"))
print(f"# {title}\n")
print(f"One local training step (ResNet-18, CIFAR shape, batch 256, bf16), every launch timed by "
      f"`ncu --metrics gpu__time_duration.sum --clock-control none` (serialised, cold caches: compare shares).\n")
print(f"* launches in the step: {len(seg)}; summed device time: {tot / 1e3:.0f} us; time inside `rlr::` (our) kernels: "
      f"{ours / 1e3:.0f} us ({100 * ours / tot:.0f} %)\n")
print("| device time (us) | share | launches | kernel |\n|---:|---:|---:|---|")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"| {v / 1e3:.1f} | {100 * v / tot:.1f} % | {c} | `{k}` |")
