"""CPU-tier checks of the bench contract: the JSON line `bench.py` prints (the last driver-format lines committed under
profiles/ by the final GPU session of a round) carries every key the driver reads, with consistent values, and the
command-line defaults are the contract's (N = 1, W >= 3).  No GPU, nothing is timed here."""
import ast
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not present")
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_default_arguments_follow_the_contract():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    defaults = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument" and node.args:
            name = getattr(node.args[0], "value", None)
            for kw in node.keywords:
                if kw.arg == "default" and isinstance(kw.value, ast.Constant):
                    defaults[name] = kw.value.value
    assert defaults["--gpus"] == 1
    assert defaults["--warmup"] >= 3 and defaults["--steps"] >= 1
    assert defaults["--impl"] == "ours"


def test_final_bench_line_has_every_key_of_the_contract():
    d = _line("r02_bench_c3_final.json")
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert baseline["metric"].startswith(d["metric"].split(" (")[0])          # BASELINE's metric, on BASELINE's config
    assert "n=10000" in d["config"]["workload"] and "model" not in d["config"]
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["data"] == "synthetic"            # BASELINE.md publishes no number for this metric
    assert d["value"] == pytest.approx(1000.0 / d["ms_per_step"], rel=1e-9)
    assert d["gpu_launches"] > 0
    # end to end through the C ABI with host buffers: copies declared, and not a repeat of the device-timed value
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]
    # roofline of the dominant kernel
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"].startswith(("TOP/s", "TFLOP/s", "GB/s"))
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0 < r["frac"] <= 1.0
    assert r["traffic"] is None or r["traffic"] > 0
    assert "traffic_source" in r and "profiles/" in r["traffic_source"]
    src = "profiles/" + r["traffic_source"].split("profiles/")[1].split(" ")[0]
    assert os.path.exists(os.path.join(ROOT, src)), src                      # the profile the constant is quoted from
    assert 0 < r["step_share"] < 1
    # CPU baseline on the box's host cores: a bounded sample, threads stated
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    # clocks sampled during the timed region; no thermal / hardware slowdown
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"]
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # parity block of the same run: north-star tolerance on the directions against the CPU oracle
    p = d["parity"]
    assert p["tol"] == 1e-8 and max(p["dir_vs_oracle"]) <= p["tol"] and max(p["kkt_residual"]) <= p["tol"]


def test_reference_arm_line_follows_the_contract():
    d = _line("r02_bench_reference_arm_final.json")
    ours = _line("r02_bench_c3_final.json")
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"] and d["config"]["workload"] == ours["config"]["workload"]
    assert d["higher_is_better"] == ours["higher_is_better"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert ours["e2e"]["value"] / d["value"] > 10.0                          # the north star's first target (>= 10 x)
