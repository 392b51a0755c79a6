mkdir -p gpurun_out/r02b
timeout 300 python scripts/diag_islands.py > gpurun_out/r02b/diag_islands.log 2>&1
timeout 400 python scripts/diag_tolerance.py > gpurun_out/r02b/diag_tolerance.log 2>&1
PROF_SKIP=102 SKIP=102 TICKS=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b/launches_t103.csv python scripts/prof_ticks.py > gpurun_out/r02b/launches_t103.log 2>&1
python scripts/launch_summary.py gpurun_out/r02b/launches_t103.csv > gpurun_out/r02b/launches_t103.summary.txt
SKIP=60 TICKS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_island_pcg" -c 4 -f -o gpurun_out/r02b/prof_islands python scripts/prof_ticks.py > gpurun_out/r02b/prof.log 2>&1
cat gpurun_out/r02b/diag_islands.log; cat gpurun_out/r02b/diag_tolerance.log; head -12 gpurun_out/r02b/launches_t103.summary.txt
