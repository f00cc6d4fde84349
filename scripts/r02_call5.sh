#!/usr/bin/env bash
# GPU call 5: CTA-per-tile sweep with explicit shared-window addressing (no S2R in the level loop): parity, trace, bench
set -u
OUT=gpurun_out/r02_c5
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_tile 300 python -m pytest tests/test_gpu_parity.py -k "tile_sweep" -q
if grep -q "rc=0 (pytest_tile)" "$OUT/steps.log"; then
for cfg in 512,1 256,1; do
  rows=${cfg%,*}; nb=${cfg#*,}
  NGSAMG_B200_TRACE_FILE=$OUT/trace_${rows}_${nb} NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=$rows,b200_tile_nbuf=$nb,log_level=info step prof_${rows}_${nb} 400 python scripts/profile_tri.py 311
  grep -a "tile sweep\|gs_tri" $OUT/prof_${rows}_${nb}.log
  python scripts/analyze_ctile_trace.py $OUT/trace_${rows}_${nb}.ctile.fwd 2>&1 | tee $OUT/trace_${rows}_${nb}_fwd.txt
  rm -f $OUT/trace_${rows}_${nb}.ctile.bwd
done
NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=512 step bench_t512 600 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline
python - "$OUT/bench_t512.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['solve_s'], d['iterations'], 'setup', d['setup_s'], 'host', d['setup_host_ms'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak']); print({k:(round(v['ms'],3), round(v['gbs'])) for k,v in d['kernels_level0'].items()}); print(d['rap']); print(d['vcycle_phases_ms']); print(d['kernel_ms_by_level'])
PY
fi
tail -n 5 "$OUT"/pytest_tile.log
cat "$OUT/steps.log"
