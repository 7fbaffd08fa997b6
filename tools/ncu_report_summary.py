#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into a short text file for profiles/:
duration, DRAM traffic (dram__bytes_read+write), L2/SM throughput, tensor-pipe activity, occupancy, issue stats,
warp-stall totals and the top stalled SASS lines (from the source page). Usage: ncu_report_summary.py REP [OUT]"""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    lines = []
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        lines.append(f"kernel: {d['Kernel Name']}")
        lines.append(f"grid {d.get('launch__grid_size')} block {d.get('launch__block_size')} "
                     f"regs {d.get('launch__registers_per_thread')} dyn smem {d.get('launch__shared_mem_per_block_dynamic')} "
                     f"{u.get('launch__shared_mem_per_block_dynamic')}")
        keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
                "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
                "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
                "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
                "sm__pipe_tmem_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
        for k in keys:
            if k in d and d[k] != "":
                lines.append(f"  {k} = {d[k]} {u.get(k, '')}")
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        data = src[2:]
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
        agg = sorted(((sum(int(r[ix[n]] or 0) for r in data), n) for n in stalls), reverse=True)
        lines.append(f"warp-stall samples: {tot}  " + "  ".join(f"{n[6:]}={100.0 * v / max(tot, 1):.1f}%" for v, n in agg[:8]))
        lines.append("top SASS lines by samples:")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
            s = int(r[ix["# Samples"]] or 0)
            st = sorted(((int(r[ix[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
            lines.append(f"  {100.0 * s / max(tot, 1):5.1f}%  {r[ix['Source']].strip()[:80]:80s} {st}")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
