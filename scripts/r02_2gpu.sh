#!/usr/bin/env bash
# 2-GPU call: NCCL transport parity (product vs the multi-rank oracle, eager and graph-captured), the two-GPU pytest, bench at N=2
set -u
OUT=gpurun_out/r02_2gpu
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
nvidia-smi -L > $OUT/gpus.txt
step nccl_parity 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_nccl_check.py
tail -n 12 $OUT/nccl_parity.log
step pytest_nccl 600 python -m pytest tests/test_gpu_parallel.py -k nccl -q
tail -n 3 $OUT/pytest_nccl.log
step bench2 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3
python - "$OUT/bench2.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('N=2', d['solve_s'], d['iterations'], 'setup', d['setup_s'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'value', d['value'], 'e2e', d['e2e']['value']); print(d['vcycle_phases_ms']); print(d['kernel_ms_by_level'][:3]); print([ (l['n'], l['gs_depth']) for l in d['config']['levels']])
PY
tail -n 5 $OUT/bench2.log | cut -c1-300
cat "$OUT/steps.log"
