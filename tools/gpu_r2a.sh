#!/bin/bash
# round-2 first visit: diagnostics + baseline tests + ncu --set full of the SHIPPED blend kernels
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python tools/parity_diag.py > $OUT/r2a_diag.log 2>&1; echo "diag exit $?" >> $OUT/r2a_diag.log
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r2a_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/r2a_pytest.log
tail -5 $OUT/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/r2a_bench.log 2> $OUT/r2a_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd" --launch-skip 14 -c 2 \
   -f -o $OUT/r2a_blend python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-graph > $OUT/r2a_ncu_run.log 2>&1
ls -la $OUT | tail -8
tail -c 1500 $OUT/r2a_diag.log
