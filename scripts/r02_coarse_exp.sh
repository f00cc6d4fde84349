#!/usr/bin/env bash
# coarse-level sweep variants on levels 1-3 of the 311^3 hierarchy
set -u
OUT=gpurun_out/r02_coarse
mkdir -p "$OUT"
for fl in "" "b200_tri_level_launch_depth=0" "b200_tri_level_launch_depth=64" "b200_tri_level_launch_depth=64,b200_tri_level_launch_rows=32768" "b200_tri_small_rows=100000,b200_tri_level_launch_depth=0"; do
  echo "=== flags: $fl"
  NGSAMG_FLAGS=$fl LEVELS=1,2,3 KERNELS=gs_tri_fwd,gs_tri_bwd timeout 400 python scripts/profile_tri.py 311 2>&1 | grep "gs_tri" | sed 's/^{[^}]*}//'
done
