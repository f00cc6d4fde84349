#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2c27
mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline > "$OUT/bench.log" 2>&1
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s2c27/bench.json'))
print('solve ms', d['ms_per_step'], 'its', d['iterations'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'])
for k,v in d['kernels_level0'].items(): print(k, round(v['ms'],3), round(v['gbs']))
for l,k in enumerate(d['kernel_ms_by_level']): print(l, k, round(sum(k.values()),3))
PY
