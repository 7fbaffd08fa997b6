#!/bin/bash
timeout 900 python -m pytest tests/test_api_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -5 | cut -c1-300
timeout 600 python - <<'PY'
import json, os, torch, bench
torch.cuda.set_device(0)
for g in ("1", "0"):
    from invertible_cd_b200 import graphs
    graphs.set_enabled(g == "1")
    r = bench.measure_small_batch("cuda:0", iters=10)
    print("graphs", g, {k: (round(v.get("ms_per_latent", v.get("ms_per_edit", v.get("ms_per_batch", -1))), 2) if "error" not in v else v) for k, v in r.items()})
PY
