#!/usr/bin/env bash
# bench with the row-major small-level sweep + ncu --set full of the level-0 streaming kernels (restrict / prolong / U-pass / L-pass)
set -u
OUT=gpurun_out/r02_s2c14
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_gpu 600 python -m pytest tests -m gpu -q
tail -n 5 "$OUT/pytest_gpu.log"
step bench 600 python bench.py --steps 3 --warmup 3 --no-multicolor --no-cpu-baseline
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s2c14/bench.json'))
print('solve ms', d['ms_per_step'], 'its', d['iterations'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'])
for l,k in enumerate(d['kernel_ms_by_level']): print(l, k, round(sum(k.values()),3))
PY
KERNELS=restrict,prolong,gs_upass,gs_lpass step ncu_stream 600 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv -c 24 -o "$OUT/stream_full" python scripts/profile_tri.py 311
ncu -i "$OUT/stream_full.ncu-rep" --page raw --csv > "$OUT/stream_full_raw.csv" 2>/dev/null
rm -f "$OUT/stream_full.ncu-rep"
cat "$OUT/steps.log"
