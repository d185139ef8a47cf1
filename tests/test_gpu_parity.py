"""GPU parity tests: CUDA path (through the C-ABI, ctypes) vs the CPU oracle on identical inputs.

Tolerance (north_star, BASELINE.json:5): <= 1e-10 relative on per-pair overlap volume, force and
torque.  The node inside/outside decisions must be bit-identical, which the exact equality of the
nodes_evaluated / nodes_inside counters checks.  PARITY UNPINNED vs the real reference (no source
in /root/reference): the oracle is this repo's own restatement, pinned by analytic KATs only.
"""
import numpy as np
import pytest

import oracle_py as O
import shpkg
from helpers import pair_rel_errors

pkg = shpkg.load()
W = pkg.workloads
TOL = 1e-10
pytestmark = pytest.mark.gpu


# pair-phase pipelines every physics test runs through: 0 = default (fused kernel below 16384 pairs, split
# cull/evaluate/reduce pipeline above), 16 = split pipeline forced (with direction-cell bound and candidate cache)
PIPELINES = [0, 16]


def both(cfg, threads=8, variant=0):
    g = pkg.ShGpu()
    o = O.Oracle(threads=threads)
    W.apply(g, cfg)
    W.apply(o, cfg)
    g.set_pair_tuning(0, 0, variant)
    return g, o


def check_forces(g, o, tol=TOL):
    g.compute_forces()
    o.compute_forces()
    cg, co = g.get_counters(), o.get_counters()
    for k in ("pair_evals", "nodes_inside"):
        assert cg[k] == co[k], (k, cg[k], co[k])
    # the GPU's conservative window / direction-cell bound visit and evaluate a SUBSET of the nodes the
    # oracle's full scan transforms and evaluates; the set of inside nodes must be identical
    assert cg["nodes_transformed"] <= co["nodes_transformed"]
    assert cg["nodes_evaluated"] <= co["nodes_evaluated"]
    e = pair_rel_errors(g.get_pairs(), o.get_pairs())
    assert e["V"] <= tol and e["F"] <= tol and e["tau"] <= tol and e["centroid"] <= tol, e
    ag, ao = g.get_atoms(), o.get_atoms()
    fs = max(1e-300, np.abs(ao["f"]).max())
    assert np.abs(ag["f"] - ao["f"]).max() <= tol * fs
    assert np.abs(ag["torque"] - ao["torque"]).max() <= tol * fs
    return e


def test_shape_tables_bit_identical():
    cfg = W.config3_packing(32, lmax=12, grid=(16, 32))
    g, o = both(cfg)
    for s in range(len(cfg["shapes"])):
        pg, po = g.shape_props(s), o.shape_props(s)
        for k in pg:
            assert np.array_equal(np.asarray(pg[k]), np.asarray(po[k])), (s, k)
        ng, no = g.nodes(s, 16 * 32), o.nodes(s, 16 * 32)
        assert np.array_equal(ng[0], no[0]) and np.array_equal(ng[1], no[1])


@pytest.mark.parametrize("variant", PIPELINES)
@pytest.mark.parametrize("exponent", [1.0, 1.5])
def test_two_particle_sweep(exponent, variant):
    """configs[0]: two SH ellipsoids (l_max=20, 32x64), sweep of separations and orientations."""
    rng = np.random.default_rng(7)
    worst = dict(V=0, F=0, tau=0)
    ncontact = 0
    for trial in range(12):
        cfg = W.config1_two_particle(seed=100 + trial, exponent=exponent)
        sep = rng.uniform(1.1, 2.05)
        dirv = rng.normal(size=3)
        dirv /= np.linalg.norm(dirv)
        cfg["x"] = np.array([-0.5 * sep * dirv, 0.5 * sep * dirv])
        g, o = both(cfg, threads=1, variant=variant)
        e = check_forces(g, o)
        ncontact += e["ncontact"]
        for k in worst:
            worst[k] = max(worst[k], e[k])
        g.close(); o.close()
    assert ncontact >= 6, "sweep produced too few contacts to be meaningful"
    print("two-particle sweep worst rel err", worst, "contacts", ncontact)


@pytest.mark.parametrize("variant", PIPELINES)
def test_packing_all_pairs_l30(variant):
    """Every pair of a periodic 8-shape l_max=30 packing snapshot (configs[2] at reduced size)."""
    cfg = W.config3_packing(500, lmax=30, grid=(48, 96))
    g, o = both(cfg, variant=variant)
    e = check_forces(g, o)
    assert e["ncontact"] > 100
    print("packing l30:", e)


@pytest.mark.parametrize("variant", PIPELINES)
def test_packing_nonperiodic(variant):
    cfg = W.packing((3, 3, 3), 20, (32, 64), nshapes=3, seed=5, periodic=False, name="np")
    g, o = both(cfg, variant=variant)
    e = check_forces(g, o)
    assert e["ncontact"] > 10


@pytest.mark.parametrize("variant", PIPELINES)
def test_wall_and_gravity_forces(variant):
    cfg = W.config2_wall(n_side=4)
    cfg["x"] = cfg["x"] - np.array([0, 0, 0.6])  # push the bottom layer into the wall
    g, o = both(cfg, variant=variant)
    check_forces(g, o)
    ag, ao = g.get_atoms(), o.get_atoms()
    assert np.abs(ao["f"][:, 2]).max() > 0
    eg, eo = g.get_energy(), o.get_energy()
    assert abs(eg["e_contact"] - eo["e_contact"]) <= 1e-10 * abs(eo["e_contact"])


@pytest.mark.parametrize("variant", PIPELINES)
def test_trajectory_two_particle_1000_steps(variant):
    """configs[0] head-on collision, 1000+ steps: GPU trajectory vs oracle trajectory."""
    cfg = W.config1_two_particle(seed=1)
    cfg["x"] = np.array([[-1.0, 0.05, 0], [1.0, -0.05, 0.02]])
    cfg["dt"] = 5e-4
    g, o = both(cfg, threads=1, variant=variant)
    g.run(1200); o.run(1200)
    ag, ao = g.get_atoms(), o.get_atoms()
    assert np.abs(ao["angmom"]).max() > 1e-6, "collision did not happen"
    for k, tol in (("x", 1e-9), ("v", 1e-8), ("quat", 1e-9), ("angmom", 1e-8)):
        assert np.abs(ag[k] - ao[k]).max() <= tol * max(1.0, np.abs(ao[k]).max()), k
    eg, eo = g.get_energy(), o.get_energy()
    for k in eg:
        assert abs(eg[k] - eo[k]) <= 1e-8 * max(1.0, abs(eo[k])), k


@pytest.mark.parametrize("variant", PIPELINES)
def test_trajectory_small_packing(variant):
    cfg = W.packing((2, 2, 2), 20, (32, 64), nshapes=2, seed=11, name="traj")
    cfg["box"] = (np.zeros(3), cfg["box"][1] * 1.0, (1, 1, 1))
    g, o = both(cfg, variant=variant)
    g.run(1000); o.run(1000)
    ag, ao = g.get_atoms(), o.get_atoms()
    for k, tol in (("x", 1e-8), ("v", 1e-7), ("quat", 1e-8), ("angmom", 1e-7)):
        assert np.abs(ag[k] - ao[k]).max() <= tol * max(1.0, np.abs(ao[k]).max()), k


@pytest.mark.parametrize("variant", PIPELINES)
def test_deterministic_bitwise(variant):
    cfg = W.config3_packing(256, lmax=20, grid=(32, 64))
    outs = []
    for _ in range(2):
        g = pkg.ShGpu(); W.apply(g, cfg); g.set_pair_tuning(0, 0, variant); g.run(20)
        outs.append(g.get_atoms()); g.close()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_neighbor_skin_does_not_change_forces():
    cfg = W.config3_packing(256, lmax=20, grid=(32, 64))
    res = []
    for skin in (0.0, 0.3):
        cfg["skin"] = skin
        g = pkg.ShGpu(); W.apply(g, cfg); g.compute_forces()
        res.append(g.get_atoms()["f"]); g.close()
    assert np.abs(res[0] - res[1]).max() <= 1e-12 * np.abs(res[0]).max()


@pytest.mark.parametrize("threads,variant", [(128, 1), (256, 1), (128, 4), (256, 4), (384, 4), (512, 4), (512, 6), (0, 0), (0, 16), (0, 18), (0, 24), (0, 26)])
def test_kernel_variants_agree_with_oracle(threads, variant):
    """variant bits: 1 = CTA-per-pair full-scan kernel; 4 = fused warp-per-pair kernel; 0 = split
    cull/evaluate/reduce pipeline (default); 2 = direction-cell bound off; 8 = candidate cache off;
    16 = force the split pipeline (small systems default to the fused kernel)."""
    cfg = W.config3_packing(256, lmax=20, grid=(32, 64))
    g, o = both(cfg)
    g.set_pair_tuning(threads, 0, variant)
    e = check_forces(g, o)
    assert e["ncontact"] > 20
    if variant == 1:      # full scan: every counter equals the oracle's
        assert g.get_counters()["nodes_transformed"] == o.get_counters()["nodes_transformed"]
    if variant in (1, 2, 6, 18, 26):  # no direction-cell bound: every bounding-sphere survivor is evaluated, as in the oracle
        assert g.get_counters()["nodes_evaluated"] == o.get_counters()["nodes_evaluated"]
    if variant in (0, 4, 16, 24):
        assert g.get_counters()["nodes_evaluated"] < o.get_counters()["nodes_evaluated"]


def test_deep_overlap_and_poles():
    """Centres almost coincident / along the body z axis: window degenerates to the full sphere."""
    rng = np.random.default_rng(11)
    for sep, dirv in ((0.05, (0, 0, 1.0)), (0.6, (0, 0, 1.0)), (0.9, (0, 0, -1.0)), (1e-7, (1.0, 0, 0)), (1.1, (1e-9, 0, 1.0))):
        cfg = W.config1_two_particle(seed=5, random_orient=False)
        dv = np.array(dirv) / np.linalg.norm(dirv)
        cfg["x"] = np.array([-0.5 * sep * dv, 0.5 * sep * dv])
        g, o = both(cfg, threads=1)
        e = check_forces(g, o)
        assert e["ncontact"] == 1
        g.close(); o.close()


@pytest.mark.parametrize("variant", PIPELINES)
def test_lmax50_tables_in_global_memory(variant):
    """configs[4] at reduced size: l_max=50, 80x160 quadrature, 8 shapes -> 255 KB of folded tables do not
    fit in shared memory, the pair kernel reads them through L1/L2 instead (SMEM_TABLES=false path)."""
    cfg = W.packing((2, 2, 2), 50, (80, 160), nshapes=8, seed=50, nn_frac=1.8, name="l50")
    g, o = both(cfg, threads=16, variant=variant)
    e = check_forces(g, o)
    assert e["ncontact"] > 10
    print("l50:", e)


def test_mixed_lmax_shapes():
    """Shapes with different l_max in one system (tables of different length, same quadrature)."""
    a1, b1 = W.ellipsoid_shape(8)
    a2, b2 = W.perturbed_shape(20, 3)
    a3, b3 = W.sphere_shape(0, 0.9)
    rng = np.random.default_rng(4)
    pos, box = W.fcc_positions((3, 3, 3), 1.75)
    n = len(pos)
    sims = []
    for mk in (lambda: pkg.ShGpu(), lambda: O.Oracle(threads=8)):
        s = mk()
        s.set_box(np.zeros(3), box, (1, 1, 1)); s.set_quadrature(32, 64)
        ids = [s.add_shape(8, a1, b1, 1.0), s.add_shape(20, a2, b2, 1.3), s.add_shape(0, a3, b3, 0.7)]
        sims.append(s)
    sid = rng.integers(0, 3, size=n).astype(np.int32)
    q = W.random_quaternions(rng, n)
    for s in sims:
        s.set_atoms(sid, pos, None, q, None)
        for i in range(3):
            for j in range(i, 3):
                s.pair_coeff(i, j, 500.0 * (1 + i + j), 1.0 + 0.25 * ((i + j) % 2))
    e = check_forces(sims[0], sims[1])
    assert e["ncontact"] > 20


def test_candidate_cache_matches_window_path_over_a_run():
    """Dynamic packing (fast particles, rotations): the candidate cache must be rebuilt before any node can
    escape it, so the per-step set of inside nodes is identical with the cache on and off, and the
    trajectories agree to rounding."""
    cfg = W.packing((3, 3, 3), 20, (32, 64), nshapes=4, seed=17, nn_frac=1.8, vel_sigma=1.5, dt=5e-4, skin=0.3, name="cache")
    rng = np.random.default_rng(3)
    cfg["angmom"] = rng.normal(0, 2.0, size=cfg["x"].shape)       # fast spins: rotation must trigger rebuilds too
    res = []
    for variant in (16, 24, 48):     # cache with remap across neighbor rebuilds / cache off / cache without remap
        g = pkg.ShGpu(); W.apply(g, cfg); g.set_pair_tuning(0, 0, variant)
        g.compute_forces(); g.reset_timers(); g.run(400)
        res.append((g.get_atoms(), g.get_counters(), dict(g.get_split_stats(), **g.get_cache_stats()))); g.close()
    (a0, c0, s0), (a1, c1, s1), (a2, c2, s2) = res
    assert s0["cache_remaps"] >= 3 and s2["cache_remaps"] == 0, (s0, s2)
    assert c2["nodes_inside"] == c1["nodes_inside"]
    for k in ("x", "v", "quat", "angmom"):
        assert np.abs(a2[k] - a1[k]).max() <= 1e-9 * max(1.0, np.abs(a1[k]).max()), k
    assert c0["nodes_inside"] == c1["nodes_inside"] and c0["pair_evals"] == c1["pair_evals"]
    assert c0["nodes_inside"] > 1000
    builds = s0["cache_builds"] + s0["cache_remaps"]
    assert builds >= 3, "cache was rebuilt %d times; the test must exercise the displacement trigger" % builds
    # (fast particles and spins exhaust the cache every few steps here; every exhaustion costs one window-path step)
    assert c0["nodes_transformed"] < 0.6 * c1["nodes_transformed"]
    for k in ("x", "v", "quat", "angmom"):
        assert np.abs(a0[k] - a1[k]).max() <= 1e-9 * max(1.0, np.abs(a1[k]).max()), k


@pytest.mark.parametrize("variant", PIPELINES)
def test_odd_quadrature_grid_and_tiny_systems(variant):
    """Quadrature sizes that are not multiples of the warp size, a single atom, and atoms far apart."""
    cfg = W.packing((2, 2, 2), 10, (9, 20), nshapes=2, seed=8, nn_frac=1.6, name="odd")
    g, o = both(cfg, variant=variant)
    e = check_forces(g, o)
    assert e["ncontact"] > 5
    g.close(); o.close()
    a, b = W.ellipsoid_shape(8)
    for xs in ([[0.0, 0, 0]], [[0.0, 0, 0], [50.0, 0, 0], [0, 70.0, 0]]):
        g = pkg.ShGpu(); g.set_quadrature(9, 20); g.add_shape(8, a, b); g.set_pair_tuning(0, 0, variant)
        g.set_atoms(np.zeros(len(xs), np.int32), np.array(xs), v=np.ones((len(xs), 3)))
        g.run(10)
        at = g.get_atoms()
        assert np.allclose(at["x"], np.array(xs) + 10 * 1e-4) and np.all(at["f"] == 0)
        assert g.get_pairs()["V"].size == 0
        g.close()


def test_snapshot_restart_roundtrip(tmp_path):
    """write_restart / read_restart: a run resumed from a snapshot continues like the uninterrupted run."""
    cfg = W.packing((3, 3, 3), 12, (16, 32), nshapes=3, seed=23, nn_frac=1.75, vel_sigma=0.4, dt=3e-4, name="snap")
    g = pkg.ShGpu(); W.apply(g, cfg); g.run(60)
    snap = tmp_path / "state.shsnap"
    g.write_snapshot(snap, step=60)
    g.run(60)
    ref = g.get_atoms(); g.close()
    g2 = pkg.ShGpu(); W.apply(g2, cfg)           # shapes / coefficients / fixes re-issued, atoms replaced by the snapshot
    assert g2.read_snapshot(snap) == 60
    g2.run(60)
    got = g2.get_atoms(); g2.close()
    for k in ("x", "v", "quat", "angmom"):
        assert np.abs(got[k] - ref[k]).max() <= 1e-11 * max(1.0, np.abs(ref[k]).max()), k
    g3 = pkg.ShGpu()
    with pytest.raises(pkg.ShGpuError):
        g3.read_snapshot(snap)                    # no shapes defined: refused
    bad = tmp_path / "bad.shsnap"; bad.write_bytes(b"not a snapshot")
    with pytest.raises(pkg.ShGpuError):
        g3.read_snapshot(bad)


def test_error_paths():
    g = pkg.ShGpu()
    with pytest.raises(pkg.ShGpuError):
        g.set_box([0, 0, 0], [1, 1, -1], [0, 0, 0])
    with pytest.raises(pkg.ShGpuError):
        g.set_atoms([0], [[0, 0, 0.0]])          # shape id out of range (no shapes yet)
    a, b = W.sphere_shape(2)
    g.add_shape(2, a, b)
    with pytest.raises(pkg.ShGpuError):
        g.set_quadrature(8, 16)                  # after add_shape
    with pytest.raises(pkg.ShGpuError):
        g.add_shape(2, -a, b)                    # r <= 0
    with pytest.raises(pkg.ShGpuError):
        g.pair_coeff(0, 0, 1.0, 0.5)             # exponent < 1
    g.set_atoms(np.zeros(0, np.int32), np.zeros((0, 3)))   # empty system is legal
    g.run(3)
    assert g.get_pairs()["V"].size == 0


def scaled_cfg(cfg, s):
    """The same packing in other length units: lengths x s, coefficients a_lm x s (r scales), velocities x s,
    stiffness x s^2 (exponent 1: F/m scales with s), so that x'(t) = s x(t) exactly."""
    c = dict(cfg)
    c["shapes"] = [(np.asarray(a) * s, np.asarray(b) * s) for (a, b) in cfg["shapes"]]
    c["x"] = np.asarray(cfg["x"]) * s
    c["v"] = np.asarray(cfg["v"]) * s
    if cfg["box"] is not None:
        lo, hi, per = cfg["box"]
        c["box"] = (np.asarray(lo) * s, np.asarray(hi) * s, per)
    c["skin"] = cfg["skin"] * s
    c["coeff"] = (cfg["coeff"][0] * s ** 2, cfg["coeff"][1])    # accelerations scale with s: the dynamics are similar
    return c


@pytest.mark.parametrize("variant", PIPELINES + [4])
@pytest.mark.parametrize("scale", [1e-3, 1e2])
def test_unit_scale_independence(scale, variant):
    """the FP32 pre-cull margins are relative to the pair's length scale, so SI-scale radii
    (1e-3) and large units (1e2) make exactly the decisions the oracle makes."""
    cfg = scaled_cfg(W.packing((3, 3, 3), 20, (32, 64), nshapes=4, seed=21, nn_frac=1.75, name="scaled"), scale)
    g, o = both(cfg, variant=variant)
    e = check_forces(g, o)
    assert e["ncontact"] > 30
    if variant == 16:
        # the pre-cull still prunes: far fewer nodes reach the exact stage than the window holds
        c = g.get_counters()
        assert c["nodes_evaluated"] < 0.2 * o.get_counters()["nodes_evaluated"]
    g.run(30); o.run(30)
    cg, co = g.get_counters(), o.get_counters()
    assert cg["nodes_inside"] == co["nodes_inside"]
    g.close(); o.close()


@pytest.mark.parametrize("knob", [("cull_wpb", 1), ("cull_wpb", 2), ("cull_wpb", 8), ("cull_lpp", 32), ("eval_pts", 2), ("eval_pts", 4),
                                  ("cache_level", 0), ("cache_level", 2), ("cube_n", 48), ("cube_n", 12)])
def test_split_pipeline_knobs(knob):
    """Every launch shape / table resolution of the split pipeline makes the oracle's decisions."""
    cfg = W.packing((3, 3, 3), 20, (32, 64), nshapes=4, seed=23, nn_frac=1.7, vel_sigma=0.8, dt=5e-4, name="knobs")
    g = pkg.ShGpu()
    if knob[0] == "cube_n":
        g.set_tuning(*knob)
    o = O.Oracle(threads=8)
    W.apply(g, cfg); W.apply(o, cfg)
    g.set_pair_tuning(0, 0, 16)
    if knob[0] != "cube_n":
        g.set_tuning(*knob)
    check_forces(g, o)
    g.run(60); o.run(60)
    cg, co = g.get_counters(), o.get_counters()
    assert cg["nodes_inside"] == co["nodes_inside"] and cg["pair_evals"] == co["pair_evals"]
    e = pair_rel_errors(g.get_pairs(), o.get_pairs())
    assert max(e["V"], e["F"], e["tau"]) <= 1e-8, e
    g.close(); o.close()


def test_full_pool_falls_back_to_the_fused_kernel():
    """A survivor pool that is too small costs time, never correctness: the pairs that do not fit are evaluated by
    the fused kernel in the same step and the host grows the pool for the next one."""
    cfg = W.packing((4, 4, 4), 20, (32, 64), nshapes=4, seed=29, nn_frac=1.6, name="deep")     # deep overlaps
    g, o = both(cfg, variant=16)
    e = check_forces(g, o)
    st = g.get_split_stats()
    assert e["ncontact"] > 100
    g.run(3); o.run(3)
    assert g.get_counters()["nodes_inside"] == o.get_counters()["nodes_inside"]
    print("deep / pool stats", st, g.get_split_stats())
    g.close(); o.close()


def _dissipative_cfg(seed=41):
    cfg = W.packing((4, 4, 3), 20, (32, 64), nshapes=4, seed=seed, periodic=True, name="dis", nn_frac=1.8, vel_sigma=0.8, dt=2e-4)
    rng = np.random.default_rng(seed)
    cfg["angmom"] = rng.normal(0, 0.5, size=cfg["angmom"].shape)
    cfg["dissipation"] = (3.0, 2.0, 0.4)       # gamma_n, gamma_t, mu
    return cfg


@pytest.mark.parametrize("variant", PIPELINES + [1, 4])
def test_dissipative_contact_terms_vs_oracle(variant):
    """viscous normal damping + Coulomb-capped friction (oracle A.5b) on a packing with translation and spin;
    per-pair F, torque (1e-10), per-atom sums, and the pressure-tensor sums."""
    cfg = _dissipative_cfg()
    g, o = both(cfg, variant=variant)
    e = check_forces(g, o)
    assert e["ncontact"] > 40
    # the dissipative part is really there: the elastic-only forces differ
    g2 = pkg.ShGpu(); c2 = dict(cfg); c2.pop("dissipation"); W.apply(g2, c2); g2.compute_forces()
    fe, fd = g2.get_atoms()["f"], g.get_atoms()["f"]
    assert np.abs(fe - fd).max() > 1e-3 * np.abs(fe).max()
    sg, so = g.get_stress(), o.get_stress()
    for k in ("virial", "kinetic"):
        assert np.abs(sg[k] - so[k]).max() <= 1e-10 * np.abs(so[k]).max(), k
    g.close(); o.close(); g2.close()


def test_dissipative_trajectory_vs_oracle():
    cfg = _dissipative_cfg(seed=43)
    g, o = both(cfg, variant=16)
    g.run(300); o.run(300)
    ag, ao = g.get_atoms(), o.get_atoms()
    assert np.abs(ag["x"] - ao["x"]).max() <= 1e-9
    assert np.abs(ag["v"] - ao["v"]).max() <= 1e-8 and np.abs(ag["angmom"] - ao["angmom"]).max() <= 1e-8
    eg, eo = g.get_energy(), o.get_energy()
    for k in eg:
        assert abs(eg[k] - eo[k]) <= 1e-8 * max(1.0, abs(eo[k])), k
    g.close(); o.close()
