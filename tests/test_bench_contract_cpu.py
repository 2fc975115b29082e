"""CPU: the committed bench line of the last tree (profiles/r03_final_bench.json) carries every key of the bench contract, and
every roofline fraction in it can be recomputed from the numbers next to it (achieved = bytes_per_unit x units / us)."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = os.path.join(ROOT, "profiles", "r03_final_bench.json")


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isfile(LINE), reason="no committed bench line")
def test_committed_bench_line_has_the_contract_keys_and_recomputable_fractions():
    d = json.load(open(LINE))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert "rays" in d["metric"] and d["unit"] == "rays/s"      # BASELINE.json: "train rays/sec fwd+bwd ..."
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    # value = rays per step / time per step
    rays = d["config"]["rays_per_gpu_per_step"] * d["n_gpus"]
    assert abs(d["value"] - rays / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    ach = r["bytes_per_unit"] * r["units_per_launch"] / (r["us_per_launch"] * 1e-6) / 1e9
    assert abs(ach - r["achieved"]) <= 1e-3 * ach and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
    for name, row in d["stage_rooflines"].items():
        if "achieved_gbs" not in row:
            continue
        a = row["bytes_per_unit"] * row["units"] / (row["us"] * 1e-6) / 1e9
        assert abs(a - row["achieved_gbs"]) <= 2e-3 * a, name
        assert abs(row["frac"] - row["achieved_gbs"] / row["peak_gbs"]) < 1e-3, name
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == rays * 3 * 3 * 4 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
    assert d["gpu_launches"] >= d["steps"] * 10


def test_rooflines_reports_the_one_launch_compositing_as_one_stage():
    b = _bench()
    stage = {"march_count": 77.0, "march_write": 13.5, "grid_encode_forward": 52.0, "field_forward": 47.5, "composite_forward": 15.5,
             "composite_backward": 2.7, "field_backward": 69.0, "grid_encode_backward": 109.0, "adam": 99.0, "pack_weights": 8.0}
    head, per = b.rooflines(stage, 269673, 12262256, {"hbm_gbs": 6552.0, "bf16_tflops_sustained": 1369.9},
                            {"stream_32MB_gbs": 15700.0, "gather_32MB_sector_gbs": 8600.0}, fused_composite=True)
    assert "composite_forward_backward" in per and "composite_backward" not in per
    assert abs(per["composite_forward_backward"]["us"] - 18.2) < 1e-6
    assert head["kernel"] == "grid_encode_backward" and head["bound"] == "l2"
    assert abs(head["frac"] - (2124.0 * 269673 / 109e-6 / 1e9) / 15700.0) < 1e-3
    assert abs(head["mlp_tensor"]["us"] - 116.5) < 1e-6
    _, per2 = b.rooflines(stage, 269673, 12262256, {"hbm_gbs": 6552.0}, None, fused_composite=False)
    assert "composite_backward" in per2 and per2["grid_encode_backward"]["bound"] == "hbm"      # no L2 probe: HBM bound
