"""CPU suite: the JSON lines bench.py printed on the B200 (committed under profiles/) carry every key of the measurement contract,
and their numbers are consistent with each other (roofline fraction, throughput vs step time, byte counts)."""
import glob
import json
import os

import pytest

import statefile as sf

PROFILES = os.path.join(sf.ROOT, "profiles")
OURS = sorted(glob.glob(os.path.join(PROFILES, "r*_bench_n*.json")))


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("path", OURS, ids=[os.path.basename(p) for p in OURS])
def test_bench_line_has_the_contract_keys(path):
    d = _line(path)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "gpu_launches"):
        assert key in d, key
    assert d["unit"] == "particle-pushes/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    # whole-job throughput = particles pushed per step on all GPUs / step time
    per_step = d["value"] * d["ms_per_step"] * 1e-3
    assert 0.9 * d["config"]["particles_per_gpu"] * d["n_gpus"] <= per_step <= 1.0001 * d["config"]["particles_per_gpu"] * d["n_gpus"]
    if "roofline" in d and d["roofline"]:
        r = d["roofline"]
        for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert key in r, key
        assert r["bound"] == "hbm" and r["unit"] == "GB/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
        # achieved = algorithmic bytes per launch / CUDA-event time of the launch
        assert abs(r["achieved"] - r["bytes_per_particle"] * r["particles_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
        if r["traffic"] is not None:      # measured DRAM bytes per launch: no wasted re-reads
            assert 1.0 <= r["traffic"] / (r["bytes_per_particle"] * r["particles_per_launch"]) < 1.1
    if d.get("e2e"):
        e = d["e2e"]
        for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
            assert key in e, key
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    if d.get("clocks"):
        c = d["clocks"]
        assert "sm_mhz" in c and "sm_max_mhz" in c and "reasons" in c
        assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_n1_line_carries_roofline_cpu_baseline_e2e_and_clocks():
    d = _line(os.path.join(PROFILES, "r1_bench_n1.json"))
    assert d["n_gpus"] == 1
    for key in ("roofline", "cpu_baseline", "e2e", "clocks"):
        assert d.get(key), key
    b = d["cpu_baseline"]
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in b, key
    assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0


def test_reference_arm_line():
    d = _line(os.path.join(PROFILES, "r1_bench_reference_arm.json"))
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    ours = _line(os.path.join(PROFILES, "r1_bench_n1.json"))
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"] and d["config"]["mesh"] == ours["config"]["mesh"]
