#!/bin/bash
# Round-end evidence on one GPU: full GPU tests, bench lines of every workload + the reference arm, the launch list of the
# c3 step and `ncu --set full` captures of the shipped kernels.  usage: bash tools/gpu_final.sh TAG
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_c3.log 2> $OUT/${TAG}_bench_c3.err; echo "c3 exit $?"
for wl in c2 c4 c5; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl > $OUT/${TAG}_bench_$wl.log 2> $OUT/${TAG}_bench_$wl.err; echo "$wl exit $?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.log 2> $OUT/${TAG}_bench_reference.err; echo "reference exit $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu-baseline --no-e2e --no-graph > $OUT/${TAG}_launches_run.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_summary.txt; head -12 $OUT/${TAG}_launch_summary.txt
K="blend_fwd|blend_bwd|tile_scatter|depth_sort|ssim_fwd|ssim_bwd|preprocess_bwd_shared|preprocess_fwd|lbs_bwd|lbs_fwd|tn_wgrad|knn_kernel|tn_heads|adam_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 24 -c 18 -f -o $OUT/${TAG}_ncu_step \
   python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu-baseline --no-e2e --no-graph > $OUT/${TAG}_ncu_step_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tn_gemm" --launch-skip 40 -c 4 -f -o $OUT/${TAG}_ncu_gemm \
   python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu-baseline --no-e2e --no-graph > $OUT/${TAG}_ncu_gemm_run.log 2>&1
ls -la $OUT/${TAG}_ncu_step.ncu-rep $OUT/${TAG}_ncu_gemm.ncu-rep
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            t = d.get("timing") or {}
            print(f.split("_bench_")[1][:-4], "value", round(d["value"], 3), "ms/step", d.get("ms_per_step"), "burst", (t.get("burst") or {}).get("value"), "e2e", d["e2e"] and round(d["e2e"]["value"], 3), "cpu", d.get("cpu_baseline") and d["cpu_baseline"]["value"], d.get("invalid"))
PY
