#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of one eager iCD step, joined (by launch
order) with the problem shapes bench.py --profile-step dumped. Usage: ncu_launch_summary.py launches.csv shapes.json"""
import csv
import json
import sys
from collections import defaultdict


def main(csv_path, shapes_path, top=40):
    with open(csv_path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per_id = {}
    order = []
    for r in csv.DictReader(lines):
        i = r["ID"]
        if i not in per_id:
            per_id[i] = {"name": r["Kernel Name"], "ns": 0.0, "dram": 0.0}
            order.append(i)
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if r.get("Metric Name") == "gpu__time_duration.sum":
            per_id[i]["ns"] = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        elif r.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            per_id[i]["dram"] += val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    rows = [(per_id[i]["name"], per_id[i]["ns"]) for i in order]
    dram = [per_id[i]["dram"] for i in order]
    shapes = json.load(open(shapes_path))
    total = sum(ns for _, ns in rows)
    fam = defaultdict(lambda: [0, 0.0, 0.0])
    for (name, ns), db in zip(rows, dram):
        key = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "").split("<")[0].split("(")[0]
        key = key.replace("void ", "").replace("icd::", "")
        if "at::native" in name or "at_cuda" in name:
            key = "torch:" + key[:50]
        fam[key][0] += 1
        fam[key][1] += ns
        fam[key][2] += db
    print(f"launches {len(rows)}  total {total / 1e6:.3f} ms (serialised, cold-cache: compare shares)")
    print("\n-- by kernel family")
    have_dram = sum(dram) > 0
    for k, (n, ns, db) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        extra = f"  dram {db / 1e6:9.1f} MB ({db / max(n, 1) / 1e6:7.3f} MB/launch, {db / max(ns, 1):7.1f} GB/s)" if have_dram else ""
        print(f"{ns / 1e6:9.3f} ms  {100 * ns / total:5.1f}%  n={n:5d}  {k}{extra}")
    # join tensor-core launches with shapes
    it = {"gemm_tc": iter([s for s in shapes if s["kind"] == "gemm_tc"]),
          "attention_tc": iter([s for s in shapes if s["kind"] == "attention_tc"])}
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for name, ns in rows:
        kind = ("gemm_tc" if ("gemm_tc_kernel" in name or "gemm2sm_tc_kernel" in name)
                else "attention_tc" if ("attention_tc_kernel" in name or "attention_smallkv_kernel" in name) else None)
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


    return fam


FAMILIES = {"gemm": ("gemm_tc_kernel", "gemm2sm_tc_kernel", "splitk_reduce_kernel"), "attention": ("attention_tc_kernel", "attention_smallkv_kernel"),
            "groupnorm": ("gn_fused_kernel", "gn_stats_kernel", "gn_apply_kernel"), "layernorm": ("layernorm_kernel",)}


if __name__ == "__main__":
    fam = main(*sys.argv[1:3])
    if len(sys.argv) > 4:
        # ncu_launch_summary.py launches.csv shapes.json <workload> <traffic.json>: per-family DRAM bytes per launch
        wl, out = sys.argv[3], sys.argv[4]
        try:
            data = json.load(open(out))
        except Exception:
            data = {}
        rec = {"source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the launches of "
                         f"one eager step ({sys.argv[1].split('/')[-1]})"}
        for name, pats in FAMILIES.items():
            n = sum(v[0] for k, v in fam.items() if any(p in k for p in pats))
            db = sum(v[2] for k, v in fam.items() if any(p in k for p in pats))
            ns = sum(v[1] for k, v in fam.items() if any(p in k for p in pats))
            if n:
                rec[f"{name}_dram_bytes_per_launch"] = db / n
                rec[f"{name}_launches"] = n
                rec[f"{name}_ms_serialised"] = ns / 1e6
        data[wl] = rec
        json.dump(data, open(out, "w"), indent=1)
