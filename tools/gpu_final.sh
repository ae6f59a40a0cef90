#!/bin/bash
# Last GPU visit of a round with little box time left: full GPU suite (no -x: every failure is wanted), smoke, bench.
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 200 python -m pytest tests -m gpu -q --durations=8 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log
timeout 200 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.log 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
tail -c 1500 $OUT/${TAG}_bench.log
