"""CPU: `bench.py --impl reference` (the reference's own block sources, or the port, on the host cores) prints the contract's JSON
line -- same metric, unit and workload string as the GPU arm, `impl`, a `cpu_baseline` describing the run and an `e2e` that repeats
the value with zero copied bytes.  The GPU arm itself needs a B200 and is exercised by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    sys.path.insert(0, ROOT)
    import bench
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["config"]["workload"] == bench.workload_text(bench.WORKLOAD)        # the GPU arm's workload, word for word
    assert line["value"] > 0 and abs(line["ms_per_step"] * 1e-3 * line["value"] - line["config"]["sample_frames_per_step"]) < 1e-6 * line["config"]["sample_frames_per_step"] + 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and line["dtype"] == "f32" and line["data"] == "synthetic"
