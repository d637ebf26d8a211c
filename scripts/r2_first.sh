#!/bin/bash
# round 2, first GPU call: parity suite, the new multi-config bench line, a long ESS window, reference arm
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2a_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/r2a_pytest.log)"
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > $OUT/r2a_bench.json 2> $OUT/r2a_bench.err; echo "bench rc=$?"; tail -3 $OUT/r2a_bench.err
timeout 600 python bench.py --steps 4 --warmup 3 --no-configs --no-cpu --ess-half-launches 1000 > $OUT/r2a_ess_long.json 2> $OUT/r2a_ess_long.err; echo "ess-long rc=$?"
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $OUT/r2a_ref.json 2> $OUT/r2a_ref.err; echo "ref rc=$?"; tail -3 $OUT/r2a_ref.err
nproc; free -g | head -2
