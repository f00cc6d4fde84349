#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2c25
mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -8
for prob in elasticity_p2:151; do
name=${prob%:*}; size=${prob#*:}
timeout 1200 python bench.py --problem $name --size $size --steps 3 --warmup 3 > "$OUT/bench_$name.log" 2> "$OUT/bench_$name.err"; echo "rc=$?"
tail -n 1 "$OUT/bench_$name.log" > "$OUT/bench_$name.json"
python - $OUT/bench_$name.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d['config']['workload'])
print('solve ms', d['ms_per_step'], 'its', d['iterations'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'setup', d['setup_s'], 'value', d['value'], 'e2e', d['e2e']['value'])
print(d['roofline'])
for l,(lv,k) in enumerate(zip(d['config']['levels'], d['kernel_ms_by_level']+[{}])): print(l, lv, k, round(sum(k.values()),3))
PY
done
