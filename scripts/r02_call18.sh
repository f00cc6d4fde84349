#!/usr/bin/env bash
# GPU call 18: pipelined solver in the tile-image kernel, 6 vs 7 CTAs per SM
set -u
OUT=gpurun_out/r02_c18
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_tile 300 python -m pytest tests/test_gpu_parity.py -k "tile_sweep" -q
tail -n 5 "$OUT"/pytest_tile.log
if grep -q "rc=0 (pytest_tile)" "$OUT/steps.log"; then
for cfg in 256,6 256,7 512,3; do
  rows=${cfg%,*}; mb=${cfg#*,}
  NGSAMG_B200_TRACE_FILE=$OUT/trace_${rows}_${mb} NGSAMG_FLAGS=b200_tile_rows=$rows,b200_tile_minb=$mb,log_level=info step prof_${rows}_${mb} 400 python scripts/profile_tri.py 311
  grep -a "tile images (f\|gs_tri" $OUT/prof_${rows}_${mb}.log
  python scripts/analyze_ctile_trace.py $OUT/trace_${rows}_${mb}.ctile.fwd 2>&1 | grep -E "hint|slab|gather|levels|tail|whole|advance" | tee $OUT/trace_${rows}_${mb}_fwd.txt
  rm -f $OUT/trace_${rows}_${mb}.ctile.*
done
fi
cat "$OUT/steps.log"
