#!/bin/bash
# One GPU-box visit.  usage: bash tools/gpu_visit.sh TAG [tests|notests] [ncu-kernel-regex|none] [extra bench args]
TAG=${1:-run}; TESTS=${2:-tests}; NCU=${3:-none}; shift 3
OUT=gpurun_out; mkdir -p $OUT
if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -4 $OUT/${TAG}_pytest.log
elif [ "$TESTS" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$TESTS" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -4 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 "$@" > $OUT/${TAG}_bench.log 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
for l in open("$OUT/${TAG}_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "burst", d.get("timing", {}).get("burst"), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), d["clocks"])
        print(json.dumps(d["roofline"]["breakdown_ms_per_step"]))
PY
if [ "$NCU" != "none" ]; then
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$NCU" --launch-skip ${NCU_SKIP:-8} -c ${NCU_COUNT:-2} \
     -f -o $OUT/${TAG}_ncu python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu-baseline --no-e2e --no-graph "$@" > $OUT/${TAG}_ncu_run.log 2>&1
  ls -la $OUT/${TAG}_ncu.ncu-rep
fi
if [ "${LAUNCH_LIST:-0}" = "1" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
     python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu-baseline --no-e2e --no-graph "$@" > $OUT/${TAG}_launches_run.log 2>&1
  python tools/launch_summary.py $OUT/${TAG}_launches.csv | head -45
fi
