"""The bench line the driver parses: checks the committed B200 lines (profiles/r0?_bench_n*.json,
written by bench.py on the GPU box) against the contract -- keys, units, and the arithmetic that ties
value, ms_per_step and the roofline object together."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0[0-9]_bench_n[0-9].json")))


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_line_follows_the_contract(path):
    # (lines measured before bench.py silenced it carry NCCL's version banner in front of the JSON)
    d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches",
              "roofline"):
        assert k in d, k
    assert d["unit"] == "GB/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = 2 * U * n_gpus / step time
    U = d["config"]["uncompressed_bytes_per_gpu"]
    assert d["value"] == pytest.approx(2 * U * d["n_gpus"] / (d["ms_per_step"] * 1e-3) / 1e9, rel=0.02)
    e = d["e2e"]
    assert e["unit"] == "GB/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                          # host copies can only cost
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6)
    assert not any(x in d["clocks"]["reasons"] for x in
                   ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"))
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    if os.path.basename(path).startswith("r02"):
        # round 2: the traffic figure names its source, the sharded path was checked against the reference
        # tool before timing, and both arms print the same config object
        assert r["traffic"] is None or r["traffic_source"]
        assert d["extra"]["sharded_parity"]["equal"] is True and d["extra"]["sharded_parity"]["round_trip"] is True
        ref = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_reference_arm.json")))
        if d["n_gpus"] == 1:
            assert ref["config"] == d["config"] and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]


@pytest.mark.parametrize("tag", ["r01", "r02"])
def test_reference_arm_line_is_committed(tag):
    d = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_bench_reference_arm.json")))
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["value"] > 0
