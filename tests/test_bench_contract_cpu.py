"""bench.py's output contract: ONE JSON line on stdout with the keys the driver reads.  Checked on the committed B200 lines
under profiles/ (what the last GPU visit printed) and, live, on the CPU reference arm (`--impl reference`)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "ocr_frames_per_sec_det_rec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])


def test_committed_b200_lines_follow_the_contract():
    for name, n in (("r01_bench_n1.json", 1), ("r01_bench_n2.json", 2)):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            text = f.read().strip()
        assert text.count("\n") == 0                      # one line
        d = json.loads(text)
        _check_common(d)
        assert d["n_gpus"] == n and d["dtype"] == "f16" and d["gpu_launches"] > 1000
        assert d["e2e"]["h2d_bytes_per_step"] == 32 * 1080 * 1920 * 3 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert d["e2e"]["value"] < d["value"]             # the end-to-end number is its own measurement, not a copy
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and not d["clocks"]["reasons"]
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
        if n == 1:
            c = d["cpu_baseline"]
            assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port" and c["cores"] >= 1


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--height", "360", "--width", "640"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-400:]
    lines = [l for l in p.stdout.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
