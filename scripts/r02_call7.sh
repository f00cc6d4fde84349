#!/usr/bin/env bash
# GPU call 7: cycle accounting of the level loop; 1 vs 3 CTAs per SM
set -u
OUT=gpurun_out/r02_c7
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
for cps in 0 1; do
  NGSAMG_B200_TRACE_FILE=$OUT/trace_$cps NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=512,b200_tri_ctas_per_sm=$cps,log_level=info KERNELS=gs_tri_fwd step prof_$cps 400 python scripts/profile_tri.py 311
  grep -a "tile sweep\|gs_tri" $OUT/prof_$cps.log
  python scripts/analyze_ctile_trace.py $OUT/trace_$cps.ctile.fwd 2>&1 | tee $OUT/trace_${cps}_fwd.txt
  rm -f $OUT/trace_$cps.ctile.fwd
done
cat "$OUT/steps.log"
