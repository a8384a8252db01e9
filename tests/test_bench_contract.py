"""CPU: the JSON line of bench.py's reference arm (the CPU implementation of the path, timed on host cores) carries the
keys of the bench contract; the GPU arm prints the same line plus roofline / clocks / gpu_launches (checked on the GPU
box by the driver)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_reference_arm_prints_one_json_line_with_the_contract_keys(kind):
    """kind = "reference": the unmodified reference staged under oracle/_ref by build() (numba-jitted: about a minute of
    compilation before the timed region); "port": the oracle restatement (--cpu-port, also the fallback)."""
    if kind == "reference" and not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "pttools")):
        pytest.skip("oracle/_ref is not staged here")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"] + (["--cpu-port"] if kind == "port" else []),
                         capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "bubble+GW spectra/sec" and d["unit"] == "spectra/s" and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["steps"] == 1 and d["value"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-9 * d["value"]
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0 and e2e["unit"] == d["unit"]
    assert abs(e2e["value"] - d["value"]) < 1e-9 * d["value"]
