#!/bin/bash
# ncu --set full of the kernels above 2 % of a tick that had no capture yet (friction sweeps, mid-cluster friction, the
# 320-thread island CG at tick 81)
OUT=gpurun_out/r03m; mkdir -p $OUT
SKIP=80 TICKS=1 timeout 120 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_gs_cluster_friction|k_gs_mid<\(bool\)0>|k_gs_mid<0>|k_island_pcg<320" -c 4 -f -o $OUT/prof_rest python scripts/prof_ticks.py > $OUT/prof.log 2>&1
tail -3 $OUT/prof.log; ls -la $OUT | awk '{print $5, $9}'
