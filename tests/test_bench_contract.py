"""The bench line contract (task statement, "Measurement"): the keys a driver parses must be present and well-formed.  Checked on
the committed line of the round's reference session (profiles/r01u_bench.json, produced by `python bench.py` on a B200) and on
bench.py's own argument handling — no GPU needed."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_recorded_bench_line_has_the_contract_keys():
    d = json.loads(open(os.path.join(REPO, "profiles", "r01u_bench.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "ms/frame" and d["higher_is_better"] is False and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert "workload" in d["config"] and "512^3" in d["config"]["workload"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"]
    assert r["actual_bound"]["frac"] > 0.5                       # the cone tracer's real bound (texture pipe) travels with it
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 100 * d["value"] and c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 1e8 and e["d2h_bytes_per_step"] > 1e7 and e["value"] >= d["value"]
    assert d["gpu_launches"] >= d["steps"] * 5
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["value"] < 4.0                                       # north_star: under 4 ms per frame on one B200
    c1 = d["c1_reference_mode"]
    assert c1["cpu_reference"]["kind"] == "reference" and c1["gpu"]["ms_per_frame"] < c1["cpu_reference"]["ms_per_frame"] / 50


def test_bench_refuses_to_run_the_product_arm_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not r.stdout.strip().startswith("{")                   # no JSON line from a path that did not run


def test_timed_oracle_copy_is_rebuilt_when_stale_and_exports_the_whole_interface(tmp_path):
    """bench.py times a -O3 -march=native copy of the oracle.  A copy left over from older sources (it travels to the GPU box with the
    snapshot) lacked a newer entry point and took the whole bench line down: the stamp covers the sources, and what comes back
    loads through the same ctypes table as the checker build."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    stamp = b.ORACLE_NATIVE_SO + ".host"
    so, flags = b.timing_oracle()
    if so == b.ORACLE_NATIVE_SO:
        open(stamp, "w").write("somebody else's build")            # as if the copy had come from another box / older sources
        before = os.path.getmtime(so)
        so2, _ = b.timing_oracle()
        assert so2 == so and open(stamp).read() != "somebody else's build" and os.path.getmtime(so) >= before
    from final184_b200 import api as A
    A.Library(so, "f184o_", product=False)                         # raises on a missing symbol
