"""bench.py's reference arm runs on the host CPU, so its JSON contract can be checked without a GPU: the keys the
driver reads, the bounded sample, and the e2e object of the reference arm (same value, no copies)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "256",
                          "--steps", "1", "--warmup", "0", *extra], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run()
    assert d["impl"] == "reference" and d["metric"] == "encode_mblocks_per_s" and d["unit"] == "Mblocks/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # default = the north-star configuration (config 3), strong-scaled
    assert "workload" in d["config"] and "DXT1" in d["config"]["workload"] and "S2TC_RANDOM_COLORS=64" in d["config"]["workload"]
    assert d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_workload():
    d = _run("--workload", "config2")
    assert "DXT5" in d["config"]["workload"] and "SRGB_MIXED" in d["config"]["workload"] and d["value"] > 0
