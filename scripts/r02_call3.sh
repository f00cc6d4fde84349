#!/usr/bin/env bash
# GPU call 3: CTA-per-tile sweep with box tiles (grid hint): parity, trace at 311^3, bench
set -u
OUT=gpurun_out/r02_c3
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_tile 300 python -m pytest tests/test_gpu_parity.py -k tile_sweep -q -x
NGSAMG_B200_TRACE_FILE=$OUT/trace311 NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=512,log_level=info step trace 600 python scripts/profile_tri.py 311
python scripts/analyze_ctile_trace.py $OUT/trace311.ctile.fwd > $OUT/trace_fwd.txt 2>&1
python scripts/analyze_ctile_trace.py $OUT/trace311.ctile.bwd > $OUT/trace_bwd.txt 2>&1
cat $OUT/trace_fwd.txt $OUT/trace_bwd.txt
NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=512,log_level=info step bench_t512 600 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline
tail -n 5 "$OUT"/pytest_tile.log; tail -n 4 "$OUT"/trace.log
for f in "$OUT"/bench_*.log; do echo "$f"; grep -a "tile s\|grid" "$f" | head -4; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['solve_s'], d['iterations'], 'setup', d['setup_s'], 'host', d['setup_host_ms'], 'vcycle', d['vcycle_ms']); print({k:(round(v['ms'],3), round(v['gbs'])) for k,v in d['kernels_level0'].items()})
PY
done
cat "$OUT/steps.log"
