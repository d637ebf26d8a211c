#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2u}; mkdir -p $OUT
RMN_TF32_TIMELINE=$OUT/${TAG}_tl_halves.txt timeout 300 python scripts/dense_timeline.py run 16384 2>&1 | tail -40
echo ---- one branch
RMN_TF32_HALVES=0 RMN_TF32_TIMELINE=$OUT/${TAG}_tl_one.txt timeout 300 python scripts/dense_timeline.py run 16384 2>&1 | tail -25
