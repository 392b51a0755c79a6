#!/bin/bash
# last build of the round: smoke, full GPU suite, the driver's bench line
OUT=gpurun_out/r03k; mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 240 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -2 $OUT/smoke.log; grep -v "^$" $OUT/pytest.log | tail -4; cut -c1-260 $OUT/bench.json
