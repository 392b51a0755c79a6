#!/bin/bash
OUT=gpurun_out/r02d; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
BODIES=64 timeout 600 python scripts/dev_s5.py > $OUT/timeline_s5_64.log 2>&1
BODIES=64 BIG_TIER=1 timeout 600 python scripts/dev_s5.py > $OUT/timeline_s5_64_big.log 2>&1
grep -v "^$" $OUT/pytest.log | grep -v "^Delaunizing\|^Creating\|^Recovering\|^Removing\|^Spreading\|^Refining\|^Optimizing\|^Writing\|^  Mesh\|^  Input\|^  Convex\|^Statistics\|^$\|^  Smallest\|^  Largest\|^  Shortest\|^  Longest\|^Output\|^Total\|Initializing\|Jettisoning\|^  Aspect\|^  Face\|^  Dihedral\|^  Smallest asp\|^  \s*[0-9.]* -" | tail -60; cat $OUT/bench.json | cut -c1-1500; tail -3 $OUT/bench.err; tail -12 $OUT/timeline_s5_64.log; tail -12 $OUT/timeline_s5_64_big.log
