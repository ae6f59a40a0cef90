#!/bin/bash
# One GPU-box visit: GPU parity tests, the bench line, the ncu launch list of the same command, and one
# `ncu --set full` capture of the hot kernels.  Run through gpurun from the repo root:
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh TAG'
# Everything lands in gpurun_out/TAG_*; copy what should be judged into profiles/.
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.log 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.log
# rows widened from SURVEY 8f: CUDA-event medians (gt fetch GB/s, ARAP fused vs torch formulation, FPS, chamfer, densify)
timeout 120 python tools/widen_bench.py > $OUT/${TAG}_widen_bench.jsonl 2> $OUT/${TAG}_widen_bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-graph \
      > $OUT/${TAG}_launches_run.log 2>&1
  # skip the probe/warm-up steps, then capture one eager step's worth of our kernels
  timeout 400 ncu --set full --clock-control none --import-source on \
      -k regex:"${NCU_KERNELS:-blend|ssim|preprocess|lbs_|emit_keys|adam|gt_fetch|arap_|fps_kernel}" --launch-skip ${NCU_SKIP:-40} -c ${NCU_COUNT:-14} \
      -f -o $OUT/${TAG}_top python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-graph \
      > $OUT/${TAG}_ncu_run.log 2>&1
  ls -la $OUT | tail -12
fi
