#!/usr/bin/env bash
# GPU call 16: tile-image kernel (kernels_itile.cuh): parity, trace at 311^3 for 256/512-row tiles
set -u
OUT=gpurun_out/r02_c27
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_tile 300 python -m pytest tests/test_gpu_parity.py -k "tile_sweep" -q
tail -n 15 "$OUT"/pytest_tile.log
if grep -q "rc=0 (pytest_tile)" "$OUT/steps.log"; then
for rows in 256; do
  NGSAMG_B200_TRACE_FILE=$OUT/trace_${rows} NGSAMG_FLAGS=b200_tile_rows=$rows,log_level=info step prof_${rows} 400 python scripts/profile_tri.py 311
  grep -a "tile images\|gs_tri" $OUT/prof_${rows}.log
  python scripts/analyze_ctile_trace.py $OUT/trace_${rows}.ctile.fwd 2>&1 | tee $OUT/trace_${rows}_fwd.txt
  rm -f $OUT/trace_${rows}.ctile.bwd
done
fi
cat "$OUT/steps.log"
