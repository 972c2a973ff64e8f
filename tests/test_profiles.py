"""CPU-only: the numbers bench.py quotes from profiles/ are reproducible from the committed raw captures."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_roofline_traffic_is_derived_from_the_committed_launch_lists(tmp_path):
    """profiles/roofline_traffic.json (what bench.py copies into roofline.traffic) == tools/summarize_profiles.py run on the
    committed ncu launch lists; the full-capture summary likewise."""
    src = tmp_path / "src"
    dst = tmp_path / "dst"
    src.mkdir(); dst.mkdir()
    for f in os.listdir(PROF):
        if f.startswith("r2_ncu_launches_") or (f.startswith("r2_ncu_full_") and f.endswith(".raw.csv")):
            shutil.copy(os.path.join(PROF, f), src / f)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_profiles.py"), str(src), str(dst)], check=True,
                   capture_output=True)
    got = json.load(open(dst / "roofline_traffic.json"))
    want = json.load(open(os.path.join(PROF, "roofline_traffic.json")))
    for wl in ("deepfm", "retrieval"):
        assert got[wl] == want[wl], wl
    assert json.load(open(dst / "r2_ncu_full_summary.json")) == json.load(open(os.path.join(PROF, "r2_ncu_full_summary.json")))


def test_bench_lines_in_profiles_carry_the_contract_keys():
    """The committed bench lines have every key of the bench contract (one JSON object each)."""
    must = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}
    for f, n in (("r2_bench_n1_final.json", 1), ("r2_bench_n2_final.json", 2), ("r2_bench_n4.json", 4), ("r2_bench_n8.json", 8)):
        j = json.load(open(os.path.join(PROF, f)))
        assert must <= set(j), (f, must - set(j))
        assert j["n_gpus"] == n and j["value"] > 0 and j["gpu_launches"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(j["roofline"])
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(j["e2e"])
        assert "workload" in j["config"] and "model" not in j["config"]
        assert not (set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
        if n == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(j["cpu_baseline"])
            assert {"cfg1_deep_hist", "cfg3_dcn_weak", "cfg5_widedeep_sharded"} <= set(j["legs"]) and "eager_gpu" in j
        else:
            assert j["parity"]["parity_ok"] is True and j["retrieval"]["sharded"]["parity_ok"] is True
