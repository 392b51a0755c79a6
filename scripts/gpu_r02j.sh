#!/bin/bash
OUT=gpurun_out/r02j; mkdir -p $OUT
SIZE=8000 TICKS=2 PIES_DEBUG_PBD=1 timeout 100 python scripts/diag_pbd_colour.py > $OUT/colour_8k.log 2>&1
SIZE=100000 TICKS=2 PIES_DEBUG_PBD=1 timeout 150 python scripts/diag_pbd_colour.py > $OUT/colour_100k.log 2>&1
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q > $OUT/pytest_kernels.log 2>&1
tail -25 $OUT/colour_8k.log | cut -c1-200; tail -25 $OUT/colour_100k.log | cut -c1-200; tail -4 $OUT/pytest_kernels.log
