#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2c12
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "sweep or wide or elasticity or vcycle or smoother" 2>&1 | tail -5
NGSAMG_B200_TRACE_FILE=$OUT/tr LEVELS=3,5 timeout 300 python scripts/profile_tri.py 151 > "$OUT/prof.log" 2>&1
tail -4 "$OUT/prof.log"
for f in $OUT/tr.l*.rm.fwd; do echo "== $f"; python scripts/analyze_rm_trace.py $f | tail -4; done
rm -f $OUT/tr.*
timeout 500 python scripts/tune_small_levels.py 151 > $OUT/tune151.log 2>&1; cat $OUT/tune151.log
