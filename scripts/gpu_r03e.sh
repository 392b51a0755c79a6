#!/bin/bash
OUT=gpurun_out/r03e; mkdir -p $OUT
timeout 600 python -m pytest tests/test_solver_gpu.py tests/test_abi.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
grep -v "^$" $OUT/pytest.log | tail -6; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r03e/bench.json') if l.startswith('{')][0])
print({k:d[k] for k in ("value","ms_per_step","phase_ms_per_step")}, "e2e", d["e2e"]["ms_per_step"])
print(json.dumps(d["roofline"])[:900])
PY
