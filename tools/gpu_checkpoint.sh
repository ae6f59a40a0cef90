#!/bin/bash
# checkpoint visit: full GPU tests + bench lines of every workload (committed under profiles/ by the caller)
TAG=${1:-ckpt}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_c3.log 2> $OUT/${TAG}_bench_c3.err; echo "c3 exit $?"
for wl in c2 c4 c5; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl > $OUT/${TAG}_bench_$wl.log 2> $OUT/${TAG}_bench_$wl.err; echo "$wl exit $?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("_bench_")[1][:-4], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "burst", round(d["timing"]["burst"]["value"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 4), d.get("invalid"), d["impl_detail"]["execution"][:60])
PY
