"""CPU: the committed bench lines (profiles/bench_r01_*gpu.json, written by bench.py on B200s) carry every key of the bench
contract, with the types and internal consistency the contract asks for, and the reference arm prints the same metric / config."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = json.load(open(os.path.join(ROOT, "BASELINE.json")))


def lines():
    return sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r01_?gpu.json")))


def test_committed_bench_lines_follow_the_contract():
    files = lines()
    assert len(files) >= 4
    for path in files:
        d = json.load(open(path))
        n = d["n_gpus"]
        assert d["metric"] == BASE["metric"] and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
        assert d["vs_baseline"] is None and not BASE["published"]              # no published number for this metric
        assert d["dtype"] == "u8" and d["data"] == "synthetic" and "model" not in d["config"] and "640x480" in d["config"]["workload"]
        assert d["warmup"] >= 3 and d["steps"] >= 1 and d["value"] > 0
        assert abs(d["value"] - n * 64 * 1e3 / d["ms_per_step"]) / d["value"] < 1e-6, path      # value = units of all ranks / max-over-ranks time
        assert "L2" in d["config"]["l2"]
        c = d["clocks"]
        assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not [r for r in c["reasons"] if "slowdown" in r or "thermal" in r]
        assert d["gpu_launches"] > 0
        e = d["e2e"]
        assert e["unit"] == "frames/s" and e["value"] > 0 and e["value"] != d["value"]
        assert e["h2d_bytes_per_step"] == 64 * 640 * 480 and e["d2h_bytes_per_step"] > 64 * 1000 * 60
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert "traffic" in r and r["peak_source"].startswith("measured")
        if n == 1:
            cb = d["cpu_baseline"]
            assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == "frames/s" and cb["sample"]
            assert r["traffic"] and 0.5 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.5       # nothing re-read
        h = d.get("hamming")
        if h:
            assert h["unit"] == "pairs/s" and h["roofline"]["bound"] == "tensor" and 0 < h["roofline"]["frac"] < 1


def test_weak_scaling_of_the_committed_lines():
    by_n = {json.load(open(p))["n_gpus"]: json.load(open(p)) for p in lines()}
    one = by_n[1]["single_lane"]["value"] if "single_lane" in by_n[1] else by_n[1]["value"]
    for n, d in by_n.items():
        per_gpu = d.get("single_lane", {}).get("value", d["value"]) / n
        assert per_gpu > 0.95 * one, (n, per_gpu, one)                          # frame shards, no collective: linear


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_same_metric_and_config():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    mine = json.load(open(os.path.join(ROOT, "profiles", "bench_r01_1gpu.json")))
    assert d["impl"] == "reference" and d["metric"] == mine["metric"] and d["unit"] == mine["unit"] and d["higher_is_better"] is True
    assert d["config"]["workload"] == mine["config"]["workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
