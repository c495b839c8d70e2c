#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list of
`bench.py` into (a) a per-kernel table for one pipeline step and (b) a small JSON with the dominant kernel's share of the
step and its DRAM traffic per launch, which bench.py reports as roofline.traffic.

usage: tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_step_summary  -> writes .md and .json
"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def load(fn):
    lines = [l for l in open(fn) if not l.startswith("==")]
    rows = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = int(row["ID"])
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("dad3d::", "")
        r = rows.setdefault(k, {"name": name, "grid": row["Grid Size"]})
        r[row["Metric Name"]] = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
    return list(rows.values())


def one_step(rows):
    """Launches from one stem kernel (s2d or SIMT conv) to the next (the steps are identical)."""
    idx = [i for i, r in enumerate(rows) if "stem_conv" in r["name"] or "stem_s2d" in r["name"]]
    if len(idx) >= 2:
        return rows[idx[0]:idx[1]]
    if idx:                                   # window started mid-step: the part before it belongs to the same (identical) step
        return rows[idx[0]:] + rows[:idx[0]]
    return rows


def main():
    src, dst = sys.argv[1], sys.argv[2]
    step = one_step(load(src))
    agg = collections.OrderedDict()
    for r in step:
        a = agg.setdefault(r["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += r.get("gpu__time_duration.sum", 0.0)
        a["rd"] += r.get("dram__bytes_read.sum", 0.0)
        a["wr"] += r.get("dram__bytes_write.sum", 0.0)
    total = sum(a["us"] for a in agg.values())
    lines = [f"# one pipeline step ({len(step)} launches, {total:.0f} us serialised under ncu; source {src})", "",
             "| kernel | launches | total us | share | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|"]
    for n, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
        lines.append(f"| `{n}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / total:.1f}% | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} |")
    lines += ["", "## every launch of the step", "", "| # | kernel | us | grid | DRAM rd MB | DRAM wr MB |", "|---|---|---|---|---|---|"]
    for i, r in enumerate(step):
        lines.append(f"| {i} | `{r['name'][:44]}` | {r.get('gpu__time_duration.sum', 0):.1f} | {r['grid']} | "
                     f"{r.get('dram__bytes_read.sum', 0) / 1e6:.1f} | {r.get('dram__bytes_write.sum', 0) / 1e6:.1f} |")
    open(dst + ".md", "w").write("\n".join(lines) + "\n")
    dom = max(agg.items(), key=lambda x: x[1]["us"])
    js = {"source": src, "step_us_serialised": total, "dominant_kernel": dom[0], "dominant_launches_per_step": dom[1]["n"],
          "dominant_share_of_step": dom[1]["us"] / total,
          "dominant_dram_bytes_per_step": dom[1]["rd"] + dom[1]["wr"],
          "dominant_dram_bytes_per_launch": (dom[1]["rd"] + dom[1]["wr"]) / max(dom[1]["n"], 1)}
    json.dump(js, open(dst + ".json", "w"), indent=1)
    print("\n".join(lines[:12]))
    print(json.dumps(js))


if __name__ == "__main__":
    main()
