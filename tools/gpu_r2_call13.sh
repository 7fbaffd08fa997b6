#!/bin/bash
# re-entry sanity: full gpu suite in one process, smoke, default bench line (timed by wall clock too)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "full suite rc=$? wall=$(( $(date +%s) - S ))s"; tail -4 gpurun_out/pytest_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
S=$(date +%s)
timeout 1200 python bench.py > gpurun_out/bench_default.out 2> gpurun_out/bench_default.err
echo "bench rc=$? wall=$(( $(date +%s) - S ))s"; tail -1 gpurun_out/bench_default.out > gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err | cut -c1-300
cut -c1-1500 gpurun_out/bench_default.json
S=$(date +%s)
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
echo "ref wall=$(( $(date +%s) - S ))s"
