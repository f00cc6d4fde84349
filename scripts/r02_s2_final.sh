#!/usr/bin/env bash
# final evidence of the session: GPU suite, default bench, ncu launch list of the bench command
set -u
OUT=gpurun_out/r02_s2_final
mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > "$OUT/bench.log" 2>&1; echo "rc=$?"
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s2_final/bench.json'))
print('solve ms', d['ms_per_step'], 'its', d['iterations'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'value', d['value'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['traffic'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 350 -c 1200 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-multicolor > $OUT/ncu.log 2>&1; echo "rc=$? (ncu)"
