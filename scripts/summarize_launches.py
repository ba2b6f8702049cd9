"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of ONE local step (delimited by two
consecutive batch-assembly launches: gather_im2col / gather_normalize).  Usage: summarize_launches.py launches.csv [title] [model text]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
what = sys.argv[3] if len(sys.argv) > 3 else "ResNet-18, CIFAR shape, batch 256, bf16"
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    try:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        rows.append((r["Kernel Name"], v, r.get("Grid Size", "")))
    except Exception:  # noqa: BLE001
        pass
idx = [i for i, r in enumerate(rows) if "gather_im2col" in r[0] or "gather_normalize" in r[0]]
seg = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows


def short(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "")
    n = re.sub(r"\(.*", "", n)
    m = re.match(r"([\w:]+)(<[^>]*>)?", n)
    return (m.group(1) + (m.group(2) or "")) if m else n[:80]


agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, _ in seg:
    k = short(n)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in agg.values())
ours = sum(v for k, (_, v) in agg.items() if k.startswith("rlr::"))
print(f"# {title}\n")
print(f"One local training step ({what}), every launch timed by `ncu --metrics gpu__time_duration.sum --clock-control none "
      f"--cache-control none` (launches serialised by the profiler; caches left warm as inside the captured graph).\n")
print(f"* launches in the step: {len(seg)}; summed device time: {tot / 1e3:.0f} us; inside `rlr::` (our) kernels: "
      f"{ours / 1e3:.0f} us ({100 * ours / tot:.1f} %); everything else: "
      f"{', '.join(sorted(k for k in agg if not k.startswith('rlr::'))) or 'nothing'}\n")
print("| device time (us) | share | launches | kernel |\n|---:|---:|---:|---|")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {v / 1e3:.1f} | {100 * v / tot:.1f} % | {c} | `{k}` |")
