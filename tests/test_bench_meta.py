"""CPU checks of what bench.py puts into its JSON line without a GPU: the SURVEY 8d arithmetic, and that the committed ncu
capture data (profiles/traffic.json) belongs to the committed kernel sources."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_flops_of_cfg2_follow_survey_8d():
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_under_test")
    c = bench.load_counts("cfg2")
    assert c["rays"] == 44086242
    flops = bench.algorithmic_flops(c)
    want = 37.0 * 228146100 + 17.0 * 900021295 + 22.0 * 231973547 + 31.0 * 50459785
    assert flops == want
    assert abs(flops / 1e9 - 30.409) < 0.001
    assert abs(flops / c["rays"] - 690.0) < 1.0  # FLOP per ray


def test_committed_capture_is_of_the_committed_sources():
    """roofline.capture_is_current_source in the bench line: the hash stored by scripts/ncu_roofline.py at capture time
    against the hash of uob_raytracer_b200/csrc now — same definition in both files, and equal for what is committed."""
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_under_test")
    roof = _load(os.path.join(ROOT, "scripts", "ncu_roofline.py"), "ncu_roofline_under_test")
    assert bench.source_hash() == roof.source_hash()
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    for key in ("cfg2", "cfg4"):
        cap = t[key]
        if cap["source_hash"] != bench.source_hash():
            # legitimate between a kernel change and the next capture: the bench line then says capture_is_current_source: false
            pytest.skip(f"profiles/traffic.json[{key}] was captured from other kernel sources: re-run scripts/gpu_round2.sh ncu and scripts/ncu_roofline.py")
        assert cap["executed_fp32_flop_per_launch"] > 0 and cap["duration_us_under_ncu"] > 0
        # the two independent counts of executed FP32 operations (SASS listing vs hardware counters) agree within 2 %
        assert abs(cap["executed_fp32_flop_per_launch"] / cap["executed_fp32_flop_hw_counters"] - 1.0) < 0.02
    assert "draw_fast_kernel<8, 1, 0, 0>" in t["cfg2"]["kernel"] and "draw_bvh_kernel<float, 8>" in t["cfg4"]["kernel"]
