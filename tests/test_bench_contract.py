"""The bench line the driver parses: keys and types of the committed output of the final code
(profiles/bench_default_r1_final.json = `python bench.py`, profiles/bench_ref_r1_final.json = `--impl reference`)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    lines = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()
    assert len(lines) == 1, "stdout of bench.py is ONE JSON line"
    return json.loads(lines[0])


def test_default_line_has_the_contract_keys():
    d = _load("bench_default_r1_final.json")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "Gbp/s" in base["metric"] and d["unit"] == "Gbp/s" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["warmup"] >= 3 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "l2" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["h2d_bytes_per_step"] <= e["host_input_bytes_per_step"]
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_line():
    d = _load("bench_ref_r1_final.json")
    assert d["impl"] == "reference" and d["unit"] == "Gbp/s" and d["value"] > 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
