"""CPU: the committed bench lines (profiles/bench_r0x_*gpu.json, written by bench.py on B200s) carry every key of the bench
contract, with the types and internal consistency the contract asks for, and the reference arm prints the same metric / config."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = json.load(open(os.path.join(ROOT, "BASELINE.json")))


def lines(tag="r01"):
    return sorted(glob.glob(os.path.join(ROOT, "profiles", f"bench_{tag}_?gpu.json")))


def latest_tag():
    return "r02" if lines("r02") else "r01"


@pytest.mark.parametrize("tag,at_least", [("r01", 4), ("r02", 2)])
def test_committed_bench_lines_follow_the_contract(tag, at_least):
    files = lines(tag)
    assert len(files) >= at_least
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
        if tag == "r02":
            # round-2 additions: sustained run, PCIe ceiling next to the end-to-end number, BASELINE config 4 in full with its self-check,
            # the other BASELINE frame shapes, ncu facts read from the committed capture
            assert d["sustained"]["seconds"] >= 2.0 and d["sustained"]["value"] > 0.9 * d["value"]
            ce = e["h2d_ceiling"]
            assert ce["value"] > 0 and e["value"] < 1.02 * ce["h2d_only"]                      # cannot beat the bus
            c4 = h["cfg4"]
            assert c4["known_answers_ok"] is True and c4["identical_on_all_ranks"] is True and c4["ranks"] == n and "10000000" in c4["workload"]
            assert c4["pairs_per_s"] > 0 and "error" not in c4
            for leg in ("config1_752x480_nf1200", "config3_1920x1080_nf2000", "config5_1280x720_nf1250"):
                assert d[leg]["value"] > 0, leg
            assert h["popc_backend"]["value"] > 0 and h["int8_backend"]["value"] > 0
            if n == 1:
                assert isinstance(d["latency"], list) and all(x["median_ms"] > 0 and x["calls"] >= 100 for x in d["latency"])
                assert r.get("ncu") or r.get("traffic")


@pytest.mark.parametrize("tag", ["r01", "r02"])
def test_weak_scaling_of_the_committed_lines(tag):
    by_n = {json.load(open(p))["n_gpus"]: json.load(open(p)) for p in lines(tag)}
    assert 1 in by_n
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
    mine = json.load(open(os.path.join(ROOT, "profiles", f"bench_{latest_tag()}_1gpu.json")))
    assert d["impl"] == "reference" and d["metric"] == mine["metric"] and d["unit"] == mine["unit"] and d["higher_is_better"] is True
    assert d["config"]["workload"] == mine["config"]["workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
