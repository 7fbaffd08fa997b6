#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests/test_f32_gpu.py -k "oracle or cfg0 or loops" -q -p no:cacheprovider -s 2>&1 | grep -E "parity|passed|failed|Error|error|assert|Mismatch|Greatest|test_" | cut -c1-400 | tail -40
