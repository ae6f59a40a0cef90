#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/r2b_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/r2b_pytest.log
tail -5 $OUT/r2b_pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd" --launch-skip 8 -c 2 \
   -f -o $OUT/r2b_blend python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-graph > $OUT/r2b_ncu_run.log 2>&1
ls -la $OUT | tail -4
