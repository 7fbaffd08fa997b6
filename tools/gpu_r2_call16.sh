#!/bin/bash
timeout 600 python tools/debug_adapters.py 2>&1 | tail -30
