#!/bin/bash
# compute-sanitizer memcheck of the small contact scene, long enough for body-body contacts (every island list, candidate
# filter, CCD, incidence tables, ordered sweeps)
OUT=gpurun_out/r03h; mkdir -p $OUT
TICKS=75 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_stack.py > $OUT/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/memcheck.log
tail -6 $OUT/memcheck.log
