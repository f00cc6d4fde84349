#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2c20
mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -s -k "digit or elasticity" 2>&1 | tail -8
NGSAMG_BENCH_VERBOSE=1 timeout 1200 python bench.py --problem elasticity_p2 --size 151 --steps 3 --warmup 3 > "$OUT/bench_p2.log" 2> "$OUT/bench_p2.err"; echo "rc=$?"
grep "bench r0" "$OUT/bench_p2.err" | tail -8
tail -n 1 "$OUT/bench_p2.log" > "$OUT/bench_p2.json"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s2c20/bench_p2.json'))
print(d['config']['workload'])
print('solve ms', d['ms_per_step'], 'its', d['iterations'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'setup', d['setup_s'], 'value', d['value'], 'e2e', d['e2e']['value'])
print(d['roofline'])
for k,v in d['kernels_level0'].items(): print(k, round(v['ms'],3), round(v['gbs']))
for l,(lv,k) in enumerate(zip(d['config']['levels'], d['kernel_ms_by_level']+[{}])): print(l, lv, k, round(sum(k.values()),3))
PY
