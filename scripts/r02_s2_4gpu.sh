#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2_4gpu
mkdir -p "$OUT"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 > $OUT/bench4.log 2>&1; echo "rc=$?"
python - "$OUT/bench4.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('N=4', d['solve_s'], d['iterations'], 'setup', d['setup_s'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'value', d['value'], 'e2e', d['e2e']['value']); print(d['vcycle_phases_ms']); print(d['config']['multi_gpu']); print([ (l['n'], l['gs_depth']) for l in d['config']['levels']])
PY
tail -2 $OUT/bench4.log | cut -c1-300
