"""CPU tests: analytic known-answer tests that pin the oracle's formulas (SURVEY §4.2 Tier A).

PARITY UNPINNED vs the real reference: /root/reference holds only README.md, so there are no golden
vectors of pair_spherharm to check against; these KATs are what pins the oracle instead."""
import numpy as np
import pytest
from scipy.special import roots_legendre, sph_harm_y

import oracle_py as O
import shpkg

W = shpkg.load().workloads


def test_a1_legendre_vs_scipy():
    worst = 0.0
    for x in np.cos(np.linspace(0.05, 3.09, 13)):
        P = O.legendre_norm(50, x)
        th = np.arccos(x)
        for l in range(0, 51, 3):
            for m in range(l + 1):
                ref = sph_harm_y(l, m, th, 0.0).real * (-1) ** m
                worst = max(worst, abs(P[l * (l + 1) // 2 + m] - ref) / max(abs(ref), 1e-3))
    assert worst < 1e-10


def test_gauss_legendre_nodes():
    for n in (2, 7, 32, 48, 80):
        x, w = O.gauss_legendre(n)
        xr, wr = roots_legendre(n)
        assert np.abs(x - xr).max() < 1e-14 and np.abs(w - wr).max() < 1e-14


@pytest.fixture(scope="module")
def ellipsoid():
    a, b = O.project_ellipsoid(20, 1.0, 0.8, 0.6)
    o = O.Oracle()
    o.set_quadrature(48, 96)
    sid = o.add_shape(20, a, b, 1.0)
    return o, sid, a, b


def test_projection_matches_numpy_workload_generator(ellipsoid):
    _, _, a, b = ellipsoid
    a2, b2 = W.ellipsoid_shape(20)
    assert np.abs(a - a2).max() < 1e-12 and np.abs(b - b2).max() < 1e-12
    assert abs(a[0] - 2.72762837157031) < 1e-12           # SURVEY KAT A3


def test_a2_a3_folded_trigfree_radius(ellipsoid):
    o, sid, a, b = ellipsoid
    rng = np.random.default_rng(0)
    d = rng.normal(size=(500, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    r = o.sh_radius(sid, d * rng.uniform(0.3, 3.0, size=(500, 1)))      # scale-invariant in |s|
    rex = 1 / np.sqrt(d[:, 0] ** 2 + (d[:, 1] / 0.8) ** 2 + (d[:, 2] / 0.6) ** 2)
    assert np.abs(r - rex).max() < 1e-7                                  # A3: SH(l<=20) vs exact ellipsoid
    th, ph = np.arccos(d[:, 2]), np.arctan2(d[:, 1], d[:, 0])
    rd = W.evaluate(20, a, b, th, ph)                                    # A2: direct Y_lm sum
    assert np.abs(r - rd).max() < 1e-13
    # poles: z = 0 must not produce NaN
    rp = o.sh_radius(sid, np.array([[0, 0, 1.0], [0, 0, -2.0]]))
    assert np.all(np.isfinite(rp)) and np.abs(rp - 0.6).max() < 1e-6


def test_a4_closed_surface_identities(ellipsoid):
    o, sid, _, _ = ellipsoid
    p, n = o.nodes(sid, 48 * 96)
    assert np.abs(n.sum(0)).max() < 1e-12
    vol = (p * n).sum() / 3
    assert abs(vol - 4 / 3 * np.pi * 0.48) < 1e-12
    pr = o.shape_props(sid)
    assert abs(pr["volume"] - vol) < 1e-12
    assert np.abs(pr["com"]).max() < 1e-12
    # solid ellipsoid inertia: m/5 (b^2+c^2) etc.
    m = pr["volume"]
    exact = sorted([m / 5 * (0.64 + 0.36), m / 5 * (1 + 0.36), m / 5 * (1 + 0.64)])
    assert np.abs(np.sort(pr["inertia"]) - exact).max() < 1e-7
    assert pr["rmax"] >= 1.0 and pr["rmin"] <= 0.6


@pytest.mark.parametrize("grid,dsep,tol", [((32, 64), 1.5, 0.05), ((80, 160), 1.9, 0.03), ((80, 160), 1.5, 0.01)])
def test_a5_sphere_sphere_lens(grid, dsep, tol):
    rng = np.random.default_rng(3)
    o = O.Oracle()
    o.set_quadrature(*grid)
    a, b = W.sphere_shape(0)
    sid = o.add_shape(0, a, b, 1.0)
    o.set_atoms([sid, sid], [[0, 0, 0], [dsep, 0, 0]], quat=rng.normal(size=(2, 4)))
    o.pair_coeff(0, 0, 2.0, 1.0)
    o.compute_forces()
    pr = o.get_pairs()
    Vex = np.pi * (4 + dsep) * (2 - dsep) ** 2 / 12
    Sex = np.pi * (1 - dsep ** 2 / 4)
    assert abs(pr["V"][0] - Vex) < tol * Vex
    F = pr["F"][0]
    assert abs(-F[0] - 2.0 * Sex) < tol * 2.0 * Sex          # F = p * pi a^2 along the line of centres
    assert np.abs(F[1:]).max() < tol * 2.0 * Sex
    assert np.abs(pr["tau_i"]).max() < 1e-12 and np.abs(pr["tau_j"]).max() < 1e-12
    at = o.get_atoms()
    assert np.abs(at["f"][0] + at["f"][1]).max() == 0.0     # antisymmetrised: momentum conserved exactly


def test_a6_two_sided_closure_ellipsoids():
    res = {}
    for grid in ((48, 96), (96, 192)):
        a, b = W.ellipsoid_shape(20)
        o = O.Oracle()
        o.set_quadrature(*grid)
        sid = o.add_shape(20, a, b, 1.0)
        q = W.random_quaternions(np.random.default_rng(5), 2)
        o.set_atoms([sid, sid], [[0, 0, 0], [1.25, 0.2, -0.1]], quat=q)
        o.compute_forces()
        res[grid] = o.get_pairs()["V"][0]
    assert res[(48, 96)] > 0 and abs(res[(48, 96)] - res[(96, 192)]) < 0.03 * res[(96, 192)]


def test_energy_conservation_two_particle():
    """Volume-based potential is energy-consistent up to quadrature discontinuities: bounded drift."""
    cfg = W.config1_two_particle(seed=1)
    cfg["x"] = np.array([[-1.0, 0.05, 0], [1.0, -0.05, 0.02]])
    cfg["dt"] = 5e-4
    o = O.Oracle()
    W.apply(o, cfg)
    e0 = o.get_energy()
    o.run(1200)
    e1 = o.get_energy()
    at = o.get_atoms()
    assert np.abs(at["angmom"]).max() > 1e-6                 # they did collide
    tot0 = e0["ke_trans"] + e0["ke_rot"] + e0["e_contact"]
    tot1 = e1["ke_trans"] + e1["ke_rot"] + e1["e_contact"]
    assert abs(tot1 - tot0) < 0.05 * tot0
    assert np.abs(np.linalg.norm(at["quat"], axis=1) - 1).max() < 1e-12


def test_wall_pushes_up_and_matches_cap_volume():
    o = O.Oracle()
    o.set_quadrature(80, 160)
    a, b = W.sphere_shape(0)
    sid = o.add_shape(0, a, b, 1.0)
    o.set_atoms([sid], [[0, 0, 0.7]])
    o.add_wall([0, 0, 0], [0, 0, 2.0], 3.0, 1.0)
    o.compute_forces()
    f = o.get_atoms()["f"][0]
    hcap = 0.3
    area = np.pi * (1 - 0.49)
    assert abs(f[2] - 3.0 * area) < 0.02 * 3.0 * area and np.abs(f[:2]).max() < 0.02
    vcap = np.pi * hcap ** 2 * (3 - hcap) / 3
    assert abs(o.get_energy()["e_contact"] - 3.0 * vcap) < 0.03 * 3.0 * vcap


def test_cell_list_equals_brute_force():
    cfg = W.packing((8, 8, 8), 4, (8, 16), nshapes=2, seed=3, name="nl")     # 2048 atoms -> cell list
    o = O.Oracle(threads=4)
    W.apply(o, cfg)
    o.compute_forces()
    pairs = o.get_pairs()
    x = cfg["x"]; L = cfg["box"][1]
    rmax = np.array([o.shape_props(s)["rmax"] for s in range(2)])[cfg["shape_id"]]
    sub = np.arange(0, len(x), 37)
    want = set()
    for i in sub:
        d = x[i] - x
        d -= L * np.rint(d / L)
        r2 = (d ** 2).sum(1)
        for j in np.nonzero(r2 < (rmax[i] + rmax + cfg["skin"]) ** 2)[0]:
            if j != i:
                want.add((min(i, j) + 1, max(i, j) + 1))
    have = set(zip(pairs["tag_i"].tolist(), pairs["tag_j"].tolist()))
    assert want <= have
    assert all((a in sub + 1 or b in sub + 1) is False or (a, b) in want for (a, b) in have if (a - 1) in sub or (b - 1) in sub)


def test_lees_edwards_image_equals_explicit_placement():
    """Lees-Edwards minimum image of the oracle: a pair that touches across the sheared y boundary feels exactly the
    force of the same two particles placed side by side without any boundary."""
    a, b = W.ellipsoid_shape(12)
    rng = np.random.default_rng(5)
    quat = W.random_quaternions(rng, 2)
    L = np.array([12.0, 10.0, 9.0])
    rate, t_steps, dt = 0.37, 25, 1e-3
    off = rate * L[1] * t_steps * dt
    xa = np.array([4.0, 9.3, 4.0])
    x_img = xa + np.array([0.25, 1.3, 0.1])            # where B's image (one box up) must sit
    xb = x_img - np.array([off, L[1], 0.0])             # B itself: one box down, displaced back by the offset
    def run(x, box, shear):
        o = O.Oracle()
        if box is not None:
            o.set_box(*box)
        if shear:
            o.set_shear(shear)
        o.set_quadrature(24, 48)
        sid = o.add_shape(12, a, b, 1.0)
        o.set_atoms(np.zeros(2, np.int32), x, np.zeros((2, 3)), quat, np.zeros((2, 3)))
        o.pair_coeff(sid, sid, 1e3, 1.0)
        o.set_neighbor(0.1); o.set_timestep(dt)
        return o
    o1 = run(np.array([xa, xb]), (np.zeros(3), L, (1, 1, 1)), rate)
    # frozen particles: advance the clock only (zero velocities, forces switched off by k = 0 during the clock steps)
    o1.pair_coeff(0, 0, 0.0, 1.0); o1.run(t_steps); o1.pair_coeff(0, 0, 1e3, 1.0)
    o1.compute_forces()
    o2 = run(np.array([xa, x_img]), None, 0.0)
    o2.compute_forces()
    f1, f2 = o1.get_atoms()["f"], o2.get_atoms()["f"]
    assert np.abs(f2).max() > 1.0                       # they do touch
    assert np.abs(f1 - f2).max() <= 1e-10 * np.abs(f2).max()
    t1, t2 = o1.get_atoms()["torque"], o2.get_atoms()["torque"]
    assert np.abs(t1 - t2).max() <= 1e-10 * np.abs(f2).max()


def test_dissipative_contact_two_spheres():
    """A.5b on two spheres: the normal damping adds gamma_n * approach speed along the line of centres, the friction is
    min(gamma_t |v_t|, mu F_n) against the tangential relative velocity, and its torque is r x F_t on both particles."""
    a, b = W.sphere_shape(4, 1.0)
    def run(gn, gt, mu, vi, Li=(0, 0, 0)):
        o = O.Oracle()
        o.set_quadrature(48, 96)
        sid = o.add_shape(4, a, b, 1.0)
        x = np.array([[-0.95, 0, 0], [0.95, 0, 0.0]])
        o.set_atoms(np.zeros(2, np.int32), x, np.array([vi, [0, 0, 0]], float), None, np.array([Li, [0, 0, 0]], float))
        o.pair_coeff(sid, sid, 1e3, 1.0)
        o.pair_dissipation(sid, sid, gn, gt, mu)
        o.set_neighbor(0.1)
        o.compute_forces()
        return o.get_atoms()
    el = run(0, 0, 0, (0.3, 0.2, 0))
    fn = -el["f"][0, 0]
    assert fn > 1.0 and abs(el["f"][0, 1]) < 1e-9 * fn
    # normal damping only: approaching at 0.3 along x
    nd = run(5.0, 0, 0, (0.3, 0.2, 0))
    assert abs(-nd["f"][0, 0] - (fn + 5.0 * 0.3)) < 1e-9 * fn
    assert np.abs(nd["f"][0] + nd["f"][1]).max() < 1e-12 * fn
    # separating fast: the normal force is clamped at zero, never attractive
    sep = run(1e6, 0, 0, (-0.3, 0, 0))
    assert np.abs(sep["f"]).max() < 1e-9 * fn
    # viscous branch of the friction: gamma_t |v_t| < mu F_n
    ft = run(0, 2.0, 0.5, (0.0, 0.2, 0))
    assert abs(ft["f"][0, 1] + 2.0 * 0.2) < 1e-9 and abs(ft["f"][1, 1] - 2.0 * 0.2) < 1e-9
    # torque = r x F_t with r = (+-0.95, 0, 0) from each centre to the centroid at the origin
    assert abs(ft["torque"][0, 2] - 0.95 * ft["f"][0, 1]) < 1e-6 and abs(ft["torque"][1, 2] - 0.95 * ft["f"][0, 1]) < 1e-6
    # Coulomb branch: gamma_t |v_t| > mu F_n
    fc = run(0, 1e6, 0.3, (0.0, 0.2, 0))
    assert abs(fc["f"][0, 1] + 0.3 * fn) < 1e-9 * fn
    # rolling contact: the surface velocity of a spinning sphere enters v_rel (w x r)
    I = 0.4 * (4.0 / 3.0 * np.pi) * 1.0     # 2/5 m R^2, unit density and radius
    sp = run(0, 2.0, 10.0, (0, 0, 0), Li=(0, 0, I * 0.7))           # w_z = 0.7 -> v at the contact = w x r = (0, 0.7 * 0.95, 0)
    assert abs(sp["f"][0, 1] + 2.0 * 0.7 * 0.95) < 2e-3


def test_stress_sums_two_particles():
    a, b = W.sphere_shape(4, 1.0)
    o = O.Oracle(); o.set_quadrature(32, 64)
    sid = o.add_shape(4, a, b, 2.0)
    x = np.array([[-0.9, 0.1, 0], [0.9, 0, 0.0]]); v = np.array([[0.5, 0, 0.2], [0, -0.3, 0]])
    o.set_atoms(np.zeros(2, np.int32), x, v, None, None)
    o.pair_coeff(sid, sid, 1e3, 1.0); o.set_neighbor(0.1); o.compute_forces()
    st, at = o.get_stress(), o.get_atoms()
    m = 2.0 * 4.0 / 3.0 * np.pi
    K = sum(m * np.outer(v[i], v[i]) for i in range(2))
    assert np.abs(st["kinetic"] - K).max() < 1e-6 * np.abs(K).max()
    Wv = np.outer(x[0] - x[1], at["f"][0])
    assert np.abs(st["virial"] - Wv).max() < 1e-12 * np.abs(Wv).max()
