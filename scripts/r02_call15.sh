#!/usr/bin/env bash
# GPU call: tile sweep on by default: full GPU suite, default bench line, ncu launch list + one full capture of k_gs_ctile
set -u
OUT=gpurun_out/r02_c28
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
step pytest_gpu 900 python -m pytest tests -m gpu -q
step bench 900 python bench.py --steps 5 --warmup 3
python - "$OUT/bench.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['solve_s'], d['iterations'], 'setup', d['setup_s'], 'host', d['setup_host_ms'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'e2e', d['e2e']['value'], 'value', d['value']); print({k:(round(v['ms'],3), round(v['gbs'])) for k,v in d['kernels_level0'].items()}); print(d['rap']); print(d['vcycle_phases_ms']); print(d['kernel_ms_by_level']); print(d['roofline']); print(d['cpu_baseline']); print(d['variant_multicolor'])
PY
KERNELS=gs_tri_fwd,gs_tri_bwd step plain_prof 400 python scripts/profile_tri.py 311 && \
KERNELS=gs_tri_fwd,gs_tri_bwd step ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:k_gs_itile -s 2 -c 2 -o $OUT/itile_n311 python scripts/profile_tri.py 311
step plain_b151 600 python bench.py --steps 1 --warmup 1 --size 151 --no-cpu-baseline --no-multicolor && \
step ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_n151.csv python bench.py --steps 1 --warmup 1 --size 151 --no-cpu-baseline --no-multicolor
tail -n 4 "$OUT"/pytest_gpu.log
cat "$OUT/steps.log"; ls -la $OUT
