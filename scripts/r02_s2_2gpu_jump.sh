#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2_2gpu_jump
mkdir -p "$OUT"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --problem elasticity_jump --size 101 --steps 3 --warmup 3 > $OUT/bench2.log 2>&1; echo "rc=$?"
python - "$OUT/bench2.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'][:200]); print('N=2', d['solve_s'], d['iterations'], 'setup', d['setup_s'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'value', d['value'], 'e2e', d['e2e']['value']); print(d['vcycle_phases_ms'])
PY
tail -4 $OUT/bench2.log | cut -c1-250
