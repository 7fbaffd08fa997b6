#!/bin/bash
ICD_LIB_PATH=$PWD/invertible_cd_b200/libicd_b200_prof.so timeout 300 python tools/attn_prof.py 2>&1 | tail -16
