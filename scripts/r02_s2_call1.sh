#!/usr/bin/env bash
# Session 2, GPU call 1: the whole GPU suite (no -x), default bench, reference arm, ncu launch list + one --set full capture of the level-0 sweep
set -u
OUT=gpurun_out/r02_s2c1
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a "$OUT/steps.log"; timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? ($name)" | tee -a "$OUT/steps.log"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
step pytest_gpu 900 python -m pytest tests -m gpu -q
tail -n 8 "$OUT/pytest_gpu.log"
step bench 600 python bench.py --steps 5 --warmup 3
tail -n 1 "$OUT/bench.log" > "$OUT/bench.json"
cut -c1-600 "$OUT/bench.json"
if grep -q "rc=0 (bench)" "$OUT/steps.log"; then
  step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 350 -c 1200 --csv --log-file "$OUT/launches.csv" \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-multicolor
  step ncu_full 600 ncu --set full --clock-control none --import-source on -k regex:'k_gs_(itile|ctile|tile|tri)$' -c 12 -o "$OUT/sweep_full" \
      python scripts/profile_tri.py 311
  ncu -i "$OUT/sweep_full.ncu-rep" --page raw --csv > "$OUT/sweep_full_raw.csv" 2>/dev/null
fi
cat "$OUT/steps.log"
