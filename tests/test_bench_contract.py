"""bench.py's contract pieces that do not need a GPU: the algorithmic-bytes formula, the roofline denominator source,
the reference arm's behaviour on non-zero ranks, and the refusal to run the product arm without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_gpu

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_formula():
    trav = {"rays": [10, 4], "nodes": [200, 80], "triangles": [100, 40], "instance_transforms": [1, 2], "shadow_misses": 3}
    closest, shadow = bench.algorithmic_bytes(trav)
    assert closest == 10 * 40 + 80 * 200 + 48 * 100 + 48 * 1
    assert shadow == 4 * 28 + 80 * 80 + 48 * 40 + 48 * 2 + 3 * 80


def test_roofline_denominator_comes_from_measured_peaks():
    peak, src = bench.measured_peaks()
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        assert src == "measured" and peak == float(json.load(open(path))["hbm_gbs"])
    else:
        assert src == "fallback" and peak == 6650.0


def test_reference_arm_is_rank0_only():
    """Under torchrun the reference arm runs on rank 0 alone: every other rank exits 0 without touching CUDA or the process group."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(has_gpu(), reason="asserts the no-GPU failure mode")
def test_product_arm_refuses_to_run_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_committed_ncu_summaries_regenerate_from_the_raw_page(tmp_path):
    """profiles/r2_all_kernels_ncu.csv and r2_trace8_ncu.csv (what bench.py reads roofline.traffic / issue_active / lanes_per_inst from)
    are exactly what tools/ncu_summarize.py makes of the committed `ncu --page raw --csv` export of the final capture."""
    import subprocess, sys
    raw = os.path.join(ROOT, "profiles", "r2_final_ncu_raw_page.csv.gz")
    if not os.path.exists(raw):
        pytest.skip("raw page not committed")
    out = str(tmp_path / "x")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summarize.py"), raw, out], stdout=subprocess.DEVNULL)
    for suffix, committed in (("_all_kernels_ncu.csv", "r2_all_kernels_ncu.csv"), ("_trace8_ncu.csv", "r2_trace8_ncu.csv")):
        assert open(out + suffix).read() == open(os.path.join(ROOT, "profiles", committed)).read(), committed
