#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2c29
mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > "$OUT/bench.log" 2>&1; echo "rc=$?"
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s2c29/bench.json'))
print('solve ms', d['ms_per_step'], 'its', d['iterations'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'value', d['value'], 'e2e', d['e2e'], 'roof', d['roofline']['frac'])
print(d['cpu_baseline']); print(d['variant_multicolor']['vcycle_ms'], d['clocks'], d['gpu_launches'])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
