"""bench.py's one-JSON-line contract: (a) the line committed from the last GPU visit of the round carries every key the
driver reads, with the metric / workload BASELINE.json names; (b) the reference arm (`--impl reference`: the oracle
port of the path on the host cores) runs here and prints the same shape.  No GPU involved."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _final_log():
    """the bench line of the latest GPU visit that was committed under profiles/ (r2*_bench.log, else round 1's)"""
    import glob
    final = os.path.join(ROOT, "profiles", "r2_final_bench_c3.log")      # the default `python bench.py` of the round's end
    if os.path.exists(final):
        return final
    logs = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2*_bench.log")), key=os.path.getmtime)
    return logs[-1] if logs else os.path.join(ROOT, "profiles", "r1p_bench.log")


FINAL_LOG = _final_log()


def _last_committed_line():
    for ln in reversed(open(FINAL_LOG).read().strip().splitlines()):
        if ln.startswith("{"):
            return FINAL_LOG, json.loads(ln)
    raise AssertionError("no JSON line found in " + FINAL_LOG)


def test_committed_bench_line_has_the_contract_keys():
    path, line = _last_committed_line()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert BASE_KEYS <= set(line), (path, BASE_KEYS - set(line))
    assert line["metric"].split(" at ")[0] in base["metric"] and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["dtype"] == "f32" and line["data"] == "synthetic" and "workload" in line["config"]
    assert "100k" in line["config"]["workload"] and "512x512" in line["config"]["workload"]      # configs[2] shard
    assert abs(line["value"] - 1e3 * line["config"]["frames_per_step_per_gpu"] * line["n_gpus"] / line["ms_per_step"]) \
        <= 1e-6 * line["value"]
    roof = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof)
    assert roof["bound"] in ("hbm", "tensor") and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    cpu = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cpu) and cpu["kind"] in ("port", "reference")
    e2e = line["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e2e)
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] != line["value"]
    assert line["gpu_launches"] > 0
    clocks = line["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(clocks)
    assert not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_runs_on_the_host():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "c2"], capture_output=True, text=True, timeout=600, env=env,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
