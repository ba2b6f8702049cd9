"""Markdown table from an `ncu -i report.ncu-rep --page raw --csv` export: one row per (kernel, grid), median over its launches.
Usage: ncu_table.py raw.csv > table.md"""
import collections
import csv
import re
import statistics
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
ix = {c: i for i, c in enumerate(h)}


def val(r, c, scale=None):
    if c not in ix or r[ix[c]] in ("", "n/a"):
        return None
    v = float(r[ix[c]].replace(",", ""))
    u = units[ix[c]]
    if scale == "us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    if scale == "MB":
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    return v


def short(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("rlr::", "")
    return re.sub(r"\(.*", "", n)


groups = collections.OrderedDict()
for r in rows[2:]:
    groups.setdefault((short(r[ix["Kernel Name"]]), r[ix["Grid Size"]].replace(" ", "")), []).append(r)
COLS = [("time us", "gpu__time_duration.sum", "us"), ("regs", "launch__registers_per_thread", None),
        ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", None),
        ("DRAM rd MB", "dram__bytes_read.sum", "MB"), ("DRAM wr MB", "dram__bytes_write.sum", "MB"),
        ("DRAM thr %", "dram__throughput.avg.pct_of_peak_sustained_elapsed", None),
        ("L2->SM MB", "l1tex__m_xbar2l1tex_read_bytes.sum", "MB"), ("L2 hit %", "lts__t_sector_hit_rate.pct", None),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
        ("issue active %", "smsp__issue_active.avg.pct", None)]
print("| kernel | grid | launches | " + " | ".join(c[0] for c in COLS) + " |")
print("|---|---|---:|" + "---:|" * len(COLS))
for (k, g), rs in groups.items():
    cells = []
    for _, c, sc in COLS:
        vs = [v for v in (val(r, c, sc) for r in rs) if v is not None]
        cells.append("-" if not vs else f"{statistics.median(vs):.1f}")
    print(f"| `{k}` | {g} | {len(rs)} | " + " | ".join(cells) + " |")
