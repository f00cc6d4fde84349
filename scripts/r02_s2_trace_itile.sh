#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2c26
mkdir -p "$OUT"
NGSAMG_B200_TRACE_FILE=$OUT/trace NGSAMG_FLAGS=log_level=info timeout 400 python scripts/profile_tri.py 311 > $OUT/prof.log 2>&1
grep -a "tile\|gs_tri" $OUT/prof.log | tail -8
python scripts/analyze_ctile_trace.py $OUT/trace.ctile.fwd 2>&1 | tee $OUT/trace_fwd.txt
rm -f $OUT/trace.ctile.*
