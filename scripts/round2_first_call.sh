#!/usr/bin/env bash
# First GPU call of the next round, batched into ONE gpurun invocation (box acquisition is charged per call):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh'
# Everything lands in gpurun_out/r02_first/ ; each step has its own timeout so that a hanging experimental kernel cannot eat the budget.
set -u
OUT=gpurun_out/r02_first
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }

# 1. the regular GPU suite (incl. the tests added at the end of round 1 that have never run on hardware)
step pytest_gpu 900 python -m pytest tests -m gpu -q
# 2. the experimental tile sweep: parity first (opt-in test), hard 120 s limit
NGSAMG_EXPERIMENTAL=1 step pytest_tile 120 python -m pytest tests/test_gpu_parity.py -k tile_sweep -q -x
# 3. baseline bench, then the same with the tile sweep on level 0 (only if the parity test passed)
step bench_base 420 python bench.py --steps 3 --warmup 3 --no-multicolor
if grep -q "rc=0 (pytest_tile)" "$OUT/steps.log"; then
  NGSAMG_FLAGS=b200_tile_sweep=1 step bench_tile 480 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline
  # 4. launch list of the tiled run (a number printed under ncu is never a bench value)
  NGSAMG_FLAGS=b200_tile_sweep=1 step ncu_launches 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
      --log-file "$OUT/launches_tile.csv" python bench.py --steps 1 --warmup 1 --no-multicolor --no-cpu-baseline --size 151
fi
tail -n 3 "$OUT"/bench_*.log 2>/dev/null | cut -c1-400
cat "$OUT/steps.log"
