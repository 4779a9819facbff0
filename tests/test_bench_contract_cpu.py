"""bench.py's JSON contract, checked on the CPU: the reference arm is run for real (it is the CPU restatement of the
reference's step on a bounded sample), the B200 arm is checked on the last line a B200 produced (profiles/, committed)."""
import glob
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "stcgan_train_images_per_sec_256x256_b16_per_gpu" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["value"] > 0 and d["ms_per_step"] > 0


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["dtype"] == "f32"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    # the stated configuration, not a sample of it: the full 16-image batch per step, --steps / --warmup honoured
    assert d["steps"] == 1 and d["warmup"] == 1 and d["config"]["global_batch"] == 16 and "full 16-image" in cb["sample"]
    assert abs(d["ms_per_step"] * d["value"] / 1e3 - 16) < 1e-6
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_last_b200_bench_line_has_every_contract_key():
    def order(f):
        b = os.path.basename(f)
        return int(b[1:3]), int(b.rsplit("_v", 1)[1].split(".")[0])
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_v*.json")), key=order)
    assert files, "no committed B200 bench line under profiles/"
    d = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    _check_common(d)
    assert d["dtype"] == "bf16" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1.2
    cb = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("port", "reference")
