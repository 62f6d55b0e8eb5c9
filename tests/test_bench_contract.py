"""The JSON lines bench.py printed on the B200 this round (kept under profiles/) carry every key the driver's contract
names, with consistent values.  CPU only: guards the contract against regressions in bench.py's output."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip("%s not recorded" % name)
    # the line itself is the last line of the file: the multi-GPU runs recorded before bench.py claimed file descriptor 1
    # for itself carry NCCL's version banner in front of it
    lines = [l for l in open(path).read().splitlines() if l.strip()]
    return json.loads(lines[-1])


@pytest.mark.parametrize("name", ["r2_bench_1gpu.json", "r2_bench_1gpu_v2.json", "r2_bench_1gpu_v3.json"])
def test_own_arm_line(name):
    d = _load(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        if k == "cpu_baseline" and d.get(k) is None and name.endswith("_v3.json"):
            continue                                                  # recorded with --no-cpu-baseline
        assert k in d, k
    assert d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - d["config"]["particles_total"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] >= 12 * d["config"]["particles_total"] and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                                   # host-fed can not beat HBM-resident
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"]
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["check"]["Nmodes_sum_ok"] is True
    c = d["cpu_baseline"]
    if c is None:                                                    # --no-cpu-baseline: no reference leg, no parity object
        return
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    p = d["check"]["parity"]                                         # same particles through the CUDA path and the reference
    assert p["nmodes_exact"] is True and p["grid_max_rel"] <= 1e-5 and p["pk_max_rel_below_nyquist"] <= 1e-5 and p["ok"] is True
    assert p["pk_max_rel"] <= p.get("tolerance_beyond_nyquist", 1e-5)


@pytest.mark.parametrize("name,own_name", [("r2_bench_1gpu_reference.json", "r2_bench_1gpu.json"),
                                           ("r2_bench_1gpu_reference_v2.json", "r2_bench_1gpu_v2.json")])
def test_reference_arm_line(name, own_name):
    d = _load(name)
    own = _load(own_name)
    assert d["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better"):
        assert d[k] == own[k], k
    # the reference arm runs a bounded sample and must say so in `config`
    assert d["config"]["sample_nside"] <= own["config"]["grid"] and d["config"]["same_config"] in (True, False)
    assert d["config"]["mas"] == own["config"]["mas"] and d["config"]["axis"] == own["config"]["axis"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")


@pytest.mark.parametrize("name,gpus", [("r2_bench_2gpu.json", 2), ("r2_bench_4gpu.json", 4), ("r2_bench_8gpu.json", 8),
                                       ("r2_bench_2gpu_v2.json", 2), ("r2_bench_4gpu_v2.json", 4), ("r2_bench_8gpu_v2.json", 8)])
def test_multi_gpu_lines(name, gpus):
    """The weak-scaling lines recorded under torchrun: whole-job value, NCCL-vs-single-GPU parity printed and green."""
    d = _load(name)
    assert d["n_gpus"] == gpus and d["scaling"] == "weak" and d["config"]["particles_total"] == gpus * 1024 ** 3
    assert abs(d["value"] - d["config"]["particles_total"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    p = d["check"]["parity"]
    assert p["ok"] is True and p["nmodes_exact"] is True and p["grid_max_rel"] <= 1e-5 and p["pk_max_rel"] <= 1e-5
    assert {c["exchange"] for c in p["cases"]} == {"grid", "particles"}
    assert d["e2e"]["h2d_bytes_per_step"] == 12 * d["config"]["particles_total"]        # totals over all ranks
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


@pytest.mark.parametrize("name", ["r2_bench_8gpu.json", "r2_bench_8gpu_v2.json"])
def test_north_star_line(name):
    """BASELINE.json's target: 2048^3 particles, PCS, 2048^3 grid, l = 0, 2, 4, 8 GPUs, under 1 s per snapshot."""
    d = _load(name)
    assert d["config"]["grid"] == 2048 and d["config"]["mas"] == ["PCS"] and d["config"]["particles_total"] == 2048 ** 3
    assert d["s_per_snapshot"] < 1.0
    assert d["roofline"]["ring_kernel_only"]["frac"] >= 0.60        # the binning kernel on the slab layout


def test_stdout_carries_only_the_json_line():
    """Whatever libraries write to file descriptor 1 (NCCL's banner, the compiled reference's prints) must not reach stdout."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench._claim_stdout(); os.write(1, b'noise from a C library\\n'); "
            "print('noise from python'); bench.emit({'a': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert "noise from a C library" in r.stderr and "noise from python" in r.stderr
