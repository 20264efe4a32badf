"""graph.py's deposit + gaussian_filter loop on the GPU against the fixture made with the
reference's own dependency (scipy) on the shipped trajectory."""
import json

import numpy as np
import pytest

from swarm_robotics_pheromones_b200.heatmap import HeatMap, gaussian_weights

pytestmark = pytest.mark.gpu


def test_weights_match_scipy():
    from scipy.ndimage import _filters
    r, w = gaussian_weights(5.0)
    ref = _filters._gaussian_kernel1d(5.0, 0, 20)
    assert r == 20 and np.array_equal(w, ref[20:]) and np.array_equal(ref, ref[::-1])


def test_graph_loop_matches_scipy_fixture(golden_dir):
    g = json.load(open(golden_dir / "graph_heatmap.json"))
    pts = json.load(open(golden_dir / "sup_replay_heatmap.json"))["positions"]
    H, W = g["shape"]
    hm = HeatMap(H, W, sigma=5.0)
    hm.accumulate(pts[:40])
    a = hm.snapshot()
    s40 = g["after_40"]
    r0, c0 = s40["crop_origin"]
    crop = np.array(s40["crop"])
    got = a[r0:r0 + 17, c0:c0 + 17]
    assert list(np.unravel_index(np.argmax(a), a.shape)) == s40["argmax"]
    assert np.array_equal(got, crop), f"max abs diff {np.abs(got - crop).max()}"  # bit-exact vs scipy
    assert a.max() == s40["max"]
    hm.accumulate(pts[40:])
    a = hm.snapshot()
    assert list(np.unravel_index(np.argmax(a), a.shape)) == g["argmax"] == [229, 152]
    assert a.max() == g["max"]
    assert abs(a.sum() - 15200.0) < 1e-6            # mass conserved by the reflect boundary
    bs = a[:497 // 71 * 71, :495 // 99 * 99].reshape(7, 71, 5, 99).sum(axis=(1, 3))
    assert np.allclose(bs, np.array(g["block_sums_7x5"]), rtol=1e-13, atol=1e-13)
    hm.close()


def test_reflect_boundary_against_scipy_random():
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(0)
    H, W = 37, 53   # smaller than the 41-tap support in one axis: multiple reflections
    hm = HeatMap(H, W, sigma=5.0)
    ref = np.zeros((H, W))
    pts = rng.uniform(0, 1, (12, 2))
    for p in pts:
        ref[abs(int(p[1] * H)), abs(int(p[0] * W))] += 20
        ref = gaussian_filter(ref, sigma=5)
    hm.accumulate(pts)
    assert np.array_equal(hm.snapshot(), ref)
    hm.close()
