#!/bin/bash
# A/B of bring-up knobs on the GPU box: for each DIMO_KNOBS value run the rasteriser parity tests and a short bench.
#   gpurun -- 'bash tools/knob_ab.sh TAG "5=0" "5=1" ...'
TAG=$1; shift
mkdir -p gpurun_out
for kv in "$@"; do
  name=$(echo "$kv" | tr '=,' '__')
  DIMO_KNOBS="$kv" timeout 300 python -m pytest tests/test_raster_gpu.py tests/test_step_gpu.py -m gpu -x -q > gpurun_out/${TAG}_${name}_pytest.log 2>&1
  echo "knobs $kv: $(tail -1 gpurun_out/${TAG}_${name}_pytest.log)"
  DIMO_KNOBS="$kv" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_${name}_bench.log 2> gpurun_out/${TAG}_${name}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${name}_bench.log").read().strip().splitlines()[-1])
    b = d["roofline"]["breakdown_ms_per_step"]
    print("knobs $kv: %.1f frames/s  %.3f ms/step  fwd %.4f bwd %.4f" % (d["value"], d["ms_per_step"], b.get("dimo_raster_blend_fwd", 0), b.get("dimo_raster_blend_bwd", 0)))
except Exception as e:
    print("knobs $kv: bench failed", e)
PY
done
