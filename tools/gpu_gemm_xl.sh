#!/bin/bash
# QKV-projection GEMM shapes (fat N, short K): heuristic choice vs forced tile shapes, one process per configuration set.
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/mb_gemm_qkv.log
import subprocess, sys
shapes = [3, 15, 18, 19, 20]
for bm, bn in [(0, 0), (128, 256), (256, 256), (128, 160), (256, 160), (128, 128)]:
    print(f"== bm={bm} bn={bn}", flush=True)
    for i in shapes:
        out = subprocess.run([sys.executable, "tools/gemm_bench.py", "--only", str(i), "--bm", str(bm), "--bn", str(bn), "--iters", "10"],
                             capture_output=True, text=True, timeout=60).stdout.strip().splitlines()
        print(out[-1] if out else "?", flush=True)
PY
