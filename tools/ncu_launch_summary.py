#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of one eager iCD step, joined (by launch
order) with the problem shapes bench.py --profile-step dumped. Usage: ncu_launch_summary.py launches.csv shapes.json"""
import csv
import json
import sys
from collections import defaultdict


def main(csv_path, shapes_path, top=40):
    rows = []
    with open(csv_path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    shapes = json.load(open(shapes_path))
    total = sum(ns for _, ns in rows)
    fam = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        key = name.split("<")[0].split("(")[0]
        key = key.replace("void ", "").replace("icd::", "")
        if "at::native" in name or "at_cuda" in name:
            key = "torch:" + key[:50]
        fam[key][0] += 1
        fam[key][1] += ns
    print(f"launches {len(rows)}  total {total / 1e6:.3f} ms (serialised, cold-cache: compare shares)")
    print("\n-- by kernel family")
    for k, (n, ns) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e6:9.3f} ms  {100 * ns / total:5.1f}%  n={n:5d}  {k}")
    # join tensor-core launches with shapes
    it = {"gemm_tc": iter([s for s in shapes if s["kind"] == "gemm_tc"]),
          "attention_tc": iter([s for s in shapes if s["kind"] == "attention_tc"])}
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for name, ns in rows:
        kind = "gemm_tc" if "gemm_tc_kernel" in name else "attention_tc" if "attention_tc_kernel" in name else None
        if kind is None:
            continue
        try:
            s = next(it[kind])
        except StopIteration:
            continue
        if kind == "gemm_tc":
            key = f"gemm M={s['M']} N={s['N']} K={s['K']} Z={s['Z']} conv={s['conv']} geglu={s['geglu']} bn={s['bn']}"
        else:
            key = f"attn B={s['B']} H={s['H']} Nq={s['Nq']} Nk={s['Nk']} D={s['D']} probs={s['probs']}"
        agg[key][0] += 1
        agg[key][1] += ns
        agg[key][2] += s["flops"]
    print("\n-- tensor-core launches by shape (TFLOP/s = algorithmic flops / ncu duration)")
    for k, (n, ns, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{ns / 1e6:9.3f} ms  {100 * ns / total:5.1f}%  n={n:4d}  {fl / ns / 1e3:8.1f} TFLOP/s  {k}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
