#!/bin/bash
# One full ncu capture of the changepoint kernel for a given (library, lanes per chain).
# Usage: scripts/profile_cp.sh <tag> <lib> <gl>
set -u
TAG=$1; LIB=$2; GL=$3; OUT=gpurun_out; mkdir -p $OUT
export RIEMANN_B200_LIB=$LIB RMN_CP_GL=$GL
python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:changepoint_kernel -s 2 -c 1 \
    -o $OUT/${TAG}_full python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu2.log 2>&1
tail -c 300 $OUT/${TAG}_plain.log
