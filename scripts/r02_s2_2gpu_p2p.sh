#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_s2_p2p
mkdir -p "$OUT"
for p2p in 1 0; do
NGSAMG_FLAGS=b200_halo_p2p=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$p2p bench.py --gpus 2 --steps 3 --warmup 3 --size 201 > $OUT/bench2_p2p$p2p.log 2>&1; echo "rc=$? (bench p2p=$p2p)"
python - "$OUT/bench2_p2p$p2p.log" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('N=2', d['solve_s'], d['iterations'], 'setup', d['setup_s'], 'vcycle', d['vcycle_ms'], d['vcycle_frac_of_peak'], 'value', d['value'], 'e2e', d['e2e']['value']); print(d['vcycle_phases_ms']); print(d['config']['multi_gpu'])
PY
tail -3 $OUT/bench2_p2p$p2p.log | cut -c1-200
done
