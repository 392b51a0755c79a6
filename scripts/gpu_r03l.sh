#!/bin/bash
# clean rebuild of the library: smoke + solver / kernel tests
OUT=gpurun_out/r03l; mkdir -p $OUT
timeout 100 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
timeout 300 python -m pytest tests/test_solver_gpu.py tests/test_kernels_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -2 $OUT/smoke.log; grep -v "^$" $OUT/pytest.log | tail -3
