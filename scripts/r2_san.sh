#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2san}; mkdir -p $OUT
timeout 600 python scripts/sanitize_small.py > $OUT/${TAG}_plain.log 2>&1; echo "plain rc=$? $(tail -1 $OUT/${TAG}_plain.log)"; grep -c " ok" $OUT/${TAG}_plain.log; grep "refused" $OUT/${TAG}_plain.log | head
which compute-sanitizer; timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/${TAG}_memcheck.log; grep -c "Invalid\|out of bounds" $OUT/${TAG}_memcheck.log
