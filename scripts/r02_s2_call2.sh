#!/usr/bin/env bash
# Session 2, GPU call 2: row-major warp-per-row sweep on the small levels (k_gs_tri_rm): parity, bench
set -u
OUT=gpurun_out/r02_s2c13
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_gpu 600 python -m pytest tests -m gpu -q
tail -n 15 "$OUT/pytest_gpu.log"
if true; then
  step bench 600 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline
  tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
  python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s2c13/bench.json'))
print('solve ms', d['ms_per_step'], 'its', d['iterations'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'])
for l,k in enumerate(d['kernel_ms_by_level']): print(l, k, round(sum(k.values()),3))
PY
fi
cat "$OUT/steps.log"
