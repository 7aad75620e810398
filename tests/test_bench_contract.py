"""The bench line contract (task statement, "Measurement"): the keys a driver parses must be present and well-formed.  Checked on
the committed lines of the round's final session (profiles/r02final_bench*.json, produced by `python bench.py` on B200s) and on
bench.py's own argument handling — no GPU needed."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(REPO, "profiles", name)).read().strip().splitlines()[-1])


def test_recorded_bench_line_has_the_contract_keys():
    d = _line("r02final_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "spec_delta", "c4_scaling"):
        assert k in d, k
    assert d["unit"] == "ms/frame" and d["higher_is_better"] is False and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert "workload" in d["config"] and "512^3" in d["config"]["workload"] and "model" not in d["config"]
    r = d["roofline"]
    # the dominant kernel is bound by the texture pipe: peak measured live by f184_microbench(0); the HBM view travels beside it
    assert r["bound"] == "texture" and "fetch/s" in r["unit"] and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"] and r["frac"] > 0.5
    assert d["other_bounds"]["trace_hbm"]["bound"] == "hbm" and 0 < d["other_bounds"]["voxelize_red"]["frac"] < 1
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 100 * d["value"] and c["sample"] and "-O3" in c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 1e8 and e["d2h_bytes_per_step"] > 1e7 and e["value"] >= d["value"]
    assert d["gpu_launches"] >= d["steps"] * 5
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["value"] < 4.0                                       # north_star: under 4 ms per frame on one B200
    c1 = d["c1_reference_mode"]
    assert c1["cpu_reference"]["kind"] == "reference" and c1["gpu"]["ms_per_frame"] < c1["cpu_reference"]["ms_per_frame"] / 50
    sd = d["spec_delta"]                                          # the Appendix-B tracer against the default one: quantified, both timed
    assert 0.05 < sd["rel_l2"] < 0.5 and sd["appendix_b_trace_solo_ms"] > sd["amended_trace_solo_ms"] > 0
    c4 = d["c4_scaling"]
    assert "1024^3" in c4["workload"] and c4["n_gpus"] == 1 and c4["ms_per_frame"] > d["value"]
    assert d["secondary_ms"]["gtao"] < 0.6                        # 4K GTAO within 0.6 ms


def test_recorded_scaling_lines():
    """the 8-GPU line of the same session: both workloads, parity against one GPU stated in the line, per-rank transfers in e2e"""
    one, d = _line("r02final_bench.json"), _line("r02final_bench_g8.json")
    assert d["n_gpus"] == 8 and d["scaling"] == "strong" and d["unit"] == one["unit"]
    assert d["parity_vs_1gpu"]["volumes"] == "bit-exact" and d["parity_vs_1gpu"]["rows"] == "bit-exact"
    c4 = d["c4_scaling"]
    assert c4["n_gpus"] == 8 and c4["parity_vs_1gpu"]["volumes"] == "bit-exact" and c4["parity_vs_1gpu"]["rows"] == "bit-exact"
    eff_c3 = one["value"] / (8 * d["value"])
    eff_c4 = one["c4_scaling"]["ms_per_frame"] / (8 * c4["ms_per_frame"])
    assert eff_c3 > 0.6 and eff_c4 > 0.5, (eff_c3, eff_c4)        # round 1: 0.42 / not measured
    assert d["e2e"]["value"] < one["e2e"]["value"] and d["gather"]["bytes_per_rank_min_max"][1] < 1e7


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
