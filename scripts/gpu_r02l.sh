#!/bin/bash
OUT=gpurun_out/r02l; mkdir -p $OUT
timeout 90 python bench.py --workload s2 --steps 6 --warmup 3 > $OUT/bench_s2.json 2> $OUT/bench_s2.err; echo "s2 exit $?"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 100 python -m pytest tests/test_solver_gpu.py -m gpu -q -k "ordered_sweep or two_box or stack" > $OUT/pytest.log 2>&1; echo "pytest exit $?"
cut -c1-1200 $OUT/bench_s2.json; tail -12 $OUT/bench_s2.err; tail -4 $OUT/pytest.log; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02l/bench.json') if l.startswith('{')][0])
print({k:d[k] for k in ("value","ms_per_step","phase_ms_per_step")}, d["e2e"]["ms_per_step"])
PY
