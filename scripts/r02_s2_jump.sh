#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2_jump
mkdir -p "$OUT"
timeout 900 python bench.py --problem elasticity_jump --size 201 --steps 3 --warmup 3 > "$OUT/bench.log" 2> "$OUT/bench.err"; echo "rc=$?"
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s2_jump/bench.json'))
print(d['config']['workload'])
print('solve ms', d['ms_per_step'], 'its', d['iterations'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'setup', d['setup_s'], 'value', d['value'], 'e2e', d['e2e']['value'])
for l,(lv,k) in enumerate(zip(d['config']['levels'], d['kernel_ms_by_level']+[{}])): print(l, lv, round(sum(k.values()),3))
PY
tail -3 "$OUT/bench.err" | cut -c1-300
