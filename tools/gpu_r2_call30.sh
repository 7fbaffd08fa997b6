#!/bin/bash
timeout 900 python -m pytest tests/test_f32_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/f32_speed.py 2>&1 | tail -8
