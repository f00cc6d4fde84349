#!/usr/bin/env bash
# GPU call 6: compact-gather CTA tile kernel (instruction-cache hypothesis): parity, trace
set -u
OUT=gpurun_out/r02_c10
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_tile 300 python -m pytest tests/test_gpu_parity.py -k "tile_sweep" -q

if grep -q "rc=0 (pytest_tile)" "$OUT/steps.log"; then
for cfg in 256,1 256,2 512,1 ; do
  rows=${cfg%,*}; nb=${cfg#*,}
  NGSAMG_B200_TRACE_FILE=$OUT/trace_${rows}_${nb} NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=$rows,b200_tile_nbuf=$nb,log_level=info step prof_${rows}_${nb} 400 python scripts/profile_tri.py 311
  grep -a "tile sweep\|gs_tri" $OUT/prof_${rows}_${nb}.log
  python scripts/analyze_ctile_trace.py $OUT/trace_${rows}_${nb}.ctile.fwd 2>&1 | tee $OUT/trace_${rows}_${nb}_fwd.txt
  rm -f $OUT/trace_${rows}_${nb}.ctile.bwd
done
fi
tail -n 5 "$OUT"/pytest_tile.log "$OUT"/pytest_adapter.log
cat "$OUT/steps.log"
