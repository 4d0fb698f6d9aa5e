"""The CPU restatement (oracle/pt_oracle.c) against golden vectors produced by the reference's own kernels on a B200
(tests/golden/make_golden.py).  IEEE libm on the CPU vs --use_fast_math on the GPU: hit ids must agree except on a
handful of edge-grazing rays, t within float noise, images within a small relative L2."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle.oracle import Oracle

GOLDEN = os.path.join(ROOT, "tests", "golden")


def cases():
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden
    return make_golden


@pytest.mark.parametrize("name", ["soup_bvh8", "cornell_bvh8", "cornell_bvh2", "atrium_bvh8"])
def test_oracle_matches_reference_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixtures are generated on the GPU box (tests/golden/make_golden.py)")
    mg = cases()
    g = np.load(path)
    c = mg.CASES[name]; w = c["size"][0]
    blob = mg.case_blob(c)
    assert int(g["triangles"]) == blob["triangles"].shape[0] and int(g["nodes"]) == blob["bvh_nodes"].size   # same scene as on the GPU box
    o = Oracle(blob, num_bounces=c["bounces"])
    hits = o.primary_hits(1)[:, :w]
    gh = g["hits"]
    assert (hits[..., 1] != gh[..., 1]).mean() < 2e-3
    same = hits[..., 1] == gh[..., 1]
    valid = same & (gh[..., 1] != 0xFFFFFFFF)
    t_cpu = hits[..., 2].view(np.float32)[valid]; t_gpu = gh[..., 2].view(np.float32)[valid]
    assert np.allclose(t_cpu, t_gpu, rtol=2e-5)
    acc = o.render(int(g["passes"]), aovs=("radiance", "albedo", "normal", "position"))
    def rel(a, b):
        a = a.astype(np.float64); b = b.astype(np.float64)
        return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30)))
    assert rel(acc["radiance"][:, :w, :3], g["radiance"]) < 2e-2
    assert rel(acc["albedo"][:, :w, :3], g["albedo"]) < 1e-2
    assert rel(acc["position"][:, :w, :3], g["position"]) < 1e-2
    counts = o.counters
    assert abs(int(counts[:8].sum()) - int(g["trace"].sum())) <= 2e-3 * int(g["trace"].sum())
    assert abs(int(counts[128:136].sum()) - int(g["shadow"].sum())) <= 2e-3 * int(g["shadow"].sum())


def test_oracle_accumulator_semantics():
    mg = cases()
    blob = mg.case_blob(mg.CASES["cornell_bvh8"])
    o = Oracle(blob, num_bounces=2)
    f1 = o.render_pass(1)["radiance"]; f2 = o.render_pass(2)["radiance"]
    acc = o.render(2)["radiance"]
    assert np.allclose(acc, f1 + (f2 - f1) / 2.0, atol=1e-6)            # pass 0 discarded, online mean of passes 1..2
