#!/bin/bash
# ncu capture of the time-sliced changepoint kernel at the bench shape (plain run first)
OUT=gpurun_out; TAG=${1:-r2bt}; mkdir -p $OUT
A="--workload changepoint --steps 2 --warmup 3 --no-cpu --no-ess --no-checks --no-configs"
timeout 300 python bench.py $A > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo plain run failed; tail -3 $OUT/${TAG}_plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $A > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:changepoint_sliced -s 3 -c 1 -o $OUT/${TAG}_full python bench.py $A > $OUT/${TAG}_ncu.log 2>&1
ls -la $OUT/${TAG}_full.ncu-rep | awk '{print $5}'
python -c "
import json; d=json.loads(open('$OUT/${TAG}_plain.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['kernel'], d['roofline']['kernel_ms_per_launch'])"
