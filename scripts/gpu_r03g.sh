#!/bin/bash
# compute-sanitizer memcheck of the small contact scene (every island list, detection, sweeps) + the solver tests
OUT=gpurun_out/r03g; mkdir -p $OUT
TICKS=40 timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_stack.py > $OUT/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/memcheck.log
timeout 400 python -m pytest tests/test_solver_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -8 $OUT/memcheck.log; grep -v "^$" $OUT/pytest.log | tail -4
