#!/bin/bash
OUT=gpurun_out/r02v; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_configs_gpu.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
PIES_B200_NO_SPLIT_ROWS=1 timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_nosplit.json 2> $OUT/bench_nosplit.err; echo "bench nosplit exit $?"
timeout 200 python scripts/diag_island_trace.py > $OUT/trace.log 2>&1
grep -v "^$" $OUT/pytest.log | tail -12; tail -8 $OUT/trace.log; python - <<'PY'
import json
for f in ("bench","bench_nosplit"):
    d=json.loads([l for l in open('gpurun_out/r02v/%s.json'%f) if l.startswith('{')][0])
    print(f, {k:d[k] for k in ("value","ms_per_step","phase_ms_per_step","pcg_iterations_per_step")}, "e2e", d["e2e"]["ms_per_step"], "ff", d["free_fall"]["ms_per_step"], "isl", d["roofline"]["avg_launch_ms"])
PY
