#!/usr/bin/env bash
# GPU call 2 of round 2: CTA-per-tile sweep (kernels_ctile.cuh): parity first, then the 311^3 bench with 512- and 256-row tiles.
set -u
OUT=gpurun_out/r02_c2
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_tile 300 python -m pytest tests/test_gpu_parity.py -k tile_sweep -q -x
if grep -q "rc=0 (pytest_tile)" "$OUT/steps.log"; then
  NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=512,log_level=info step bench_t512 600 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline
  NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=256,log_level=info step bench_t256 600 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline
  NGSAMG_FLAGS=b200_tile_sweep=1,b200_tile_rows=512,b200_tri_sleep_ns=0 step bench_t512_nosleep 600 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline
fi
tail -n 5 "$OUT"/pytest_tile.log
for f in "$OUT"/bench_*.log; do echo "$f"; grep -a "tile s" "$f" | head -4; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['solve_s'], d['iterations'], 'setup', d['setup_s'], 'vcycle', d['vcycle_ms']); print({k:(round(v['ms'],3), round(v['gbs'])) for k,v in d['kernels_level0'].items()})
PY
done
cat "$OUT/steps.log"
