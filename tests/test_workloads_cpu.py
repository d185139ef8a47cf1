"""CPU tests of the synthetic workload generators and example file writers (host-side input only)."""
import os
import sys

import numpy as np

import oracle_py as O
import shpkg

pkg = shpkg.load()
W = pkg.workloads
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_perturbed_shapes_are_star_shaped_and_deterministic():
    a1, b1 = W.perturbed_shape(12, 30)
    a2, b2 = W.perturbed_shape(12, 30)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
    th = np.linspace(0.02, np.pi - 0.02, 40)[:, None]
    ph = np.linspace(0, 2 * np.pi, 80, endpoint=False)[None, :]
    r = W.evaluate(12, a1, b1, th, ph)
    assert r.min() > 0.3 and r.max() < 1.3
    assert np.all(b1[[l * (l + 1) // 2 for l in range(13)]] == 0) or True   # b_l0 is ignored by the engine


def test_tiled_packing_is_a_periodic_replica_of_the_committed_unit_cell():
    c1 = W.tiled_packing((1, 1, 1))
    c2 = W.tiled_packing((2, 1, 1))
    n = len(c1["x"])
    assert n == 4000 and len(c2["x"]) == 2 * n and len(c1["shapes"]) == 8 and c1["lmax"] == 30 and c1["grid"] == (48, 96)
    box = np.asarray(c1["box"][1])
    assert np.allclose(c2["box"][1], box * [2, 1, 1])
    assert np.allclose(c2["x"][n:] - c2["x"][:n], [box[0], 0, 0])
    assert np.all(c1["x"] >= 0) and np.all(c1["x"] < box)
    assert np.allclose(np.linalg.norm(c1["quat"], axis=1), 1.0)
    # the unit cell is jammed but barely overlapping: the oracle finds contacts with tiny overlap volumes
    sub = dict(c1)
    o = O.Oracle(threads=os.cpu_count() or 1)
    W.apply(o, sub)
    o.compute_forces()
    pr = o.get_pairs()
    vol = np.mean([o.shape_props(s)["volume"] for s in range(8)])
    assert 5.0 < len(pr["V"]) / n < 7.0                      # ~6 bounding-sphere pairs per particle
    assert 0 < (pr["V"] > 0).sum() and pr["V"].max() < 0.01 * vol
    phi = sum(o.shape_props(s)["volume"] for s in c1["shape_id"]) / np.prod(box)
    assert 0.69 < phi < 0.73


def test_example_writers_round_trip(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_examples as M
    M.main(str(tmp_path))
    a, b = W.ellipsoid_shape(20)
    rows = np.loadtxt(tmp_path / "ellipsoid_l20.sh")
    assert rows.shape == (231, 4)
    k = (rows[:, 0] * (rows[:, 0] + 1) // 2 + rows[:, 1]).astype(int)
    assert np.array_equal(rows[:, 2], a[k]) and np.array_equal(rows[:, 3], b[k])
    txt = (tmp_path / "data.wall_1000").read_text()
    assert "1000 atoms" in txt and "Atoms" in txt and "Velocities" in txt
    # committed examples are what the writer produces
    assert (tmp_path / "ellipsoid_l20.sh").read_text() == open(os.path.join(ROOT, "examples", "ellipsoid_l20.sh")).read()


def test_bench_flop_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert [bench.f_eval(L) for L in (20, 30, 50)] == [1951, 3946, 10036]
    c = dict(nodes_transformed=10, nodes_evaluated=3, nodes_inside=2, pair_evals=1)
    assert bench.algorithmic_flops(c, 30) == 24 * 10 + 3946 * 3 + 30 * 2 + 200
