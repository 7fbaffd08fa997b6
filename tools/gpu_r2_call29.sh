#!/bin/bash
timeout 900 python -m pytest tests/test_adapters_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -5 | cut -c1-300
