"""CPU tests of the PROVEN tables behind the pair phase's work reduction: the product's host-side table
builder (csrc/shape_tables.cpp, compiled into a test harness) against the oracle's exact radius evaluation.

The tables are not sample-plus-heuristic-pad any more; they follow from an inequality:
  * H1 = sum_l l B_l and H2 = sum_l l^2 B_l with B_l = sqrt((2l+1)/4pi) |c_l|_2 bound |dr/dt| and |d2r/dt2| along every great
    circle (Cauchy-Schwarz + addition theorem for |f_l| <= B_l, Bernstein's inequality for the degree-l trigonometric
    polynomial f_l restricted to a great circle);
  * on a gnomonic grid of spacing D every direction in a grid square has max(corners) + H2 D^2/4 >= r >= min(corners) - H2 D^2/4.
The tests (1) recompute H1/H2 independently from the coefficients, (2) check the derivative bounds against numerical
derivatives along random great circles, (3) recompute every table entry from the formula with samples taken from the
ORACLE's radius evaluation, and only then (4) sample: ub2[cell(d)] >= r(d)^2 >= lb2[cell(d)] with the FP32 cell index of the
kernels, the candidate-cache tables dominate every direction a node can drift to, rmin <= r <= rmax."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O
import shpkg

W = shpkg.load().workloads
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("sth") / "libshtables_test.so"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread", "-o", str(out),
                           os.path.join(ROOT, "tests", "shape_tables_harness.cpp"),
                           os.path.join(ROOT, "lammps-spherharm_b200", "csrc", "shape_tables.cpp")])
    lib = C.CDLL(str(out))
    lib.sth_build.restype = C.c_void_p
    lib.sth_build.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.sth_free.argtypes = [C.c_void_p]
    lib.sth_cube_n.argtypes = [C.c_void_p]
    lib.sth_scalars.argtypes = [C.c_void_p, C.c_void_p]
    lib.sth_cube.argtypes = [C.c_void_p] * 4
    lib.sth_nodes.argtypes = [C.c_void_p] * 3
    return lib


def cube_cell(s, cn):
    """FP32 replica of cube_cell() in csrc/pair_split_kernels.cuh."""
    f = s.astype(np.float32)
    a = np.abs(f)
    major = np.where((a[:, 0] >= a[:, 1]) & (a[:, 0] >= a[:, 2]), 0, np.where(a[:, 1] >= a[:, 2], 1, 2))
    idx = np.arange(len(f))
    fm = f[idx, major]
    face = 2 * major + (fm <= 0)
    uu = np.where(major == 0, f[:, 1], f[:, 0])
    vv = np.where(major == 2, f[:, 1], f[:, 2])
    im = np.float32(1.0) / np.maximum(a[idx, major], np.float32(1e-30))
    hn = np.float32(0.5 * cn)
    iu = np.clip(((uu * im + np.float32(1.0)) * hn).astype(np.int32), 0, cn - 1)
    iv = np.clip(((vv * im + np.float32(1.0)) * hn).astype(np.int32), 0, cn - 1)
    return (face * cn + iu) * cn + iv


def face_dirs(f, u, v):
    major = f // 2
    d = np.zeros(u.shape + (3,))
    d[..., major] = -1.0 if f % 2 else 1.0
    d[..., 1 if major == 0 else 0] = u
    d[..., 1 if major == 2 else 2] = v
    return d


def h_bounds(lmax, a, b):
    """H1, H2 recomputed from the coefficients (independent of the C++ builder)."""
    h1 = h2 = 0.0
    for l in range(1, lmax + 1):
        k0 = l * (l + 1) // 2
        n2 = a[k0] ** 2 + 0.5 * float(np.sum(a[k0 + 1:k0 + l + 1] ** 2 + b[k0 + 1:k0 + l + 1] ** 2))
        B = np.sqrt((2 * l + 1) / (4 * np.pi)) * np.sqrt(n2)
        h1 += l * B
        h2 += l * l * B
    return h1, h2


SHAPES = [("ellipsoid_l20", 20, lambda: W.ellipsoid_shape(20)), ("perturbed_l12", 12, lambda: W.perturbed_shape(12, 31)),
          ("perturbed_l30", 30, lambda: W.perturbed_shape(30, 33)), ("flat_l16", 16, lambda: W.project(16, W.ellipsoid_radius(1.0, 0.9, 0.45)))]


@pytest.mark.parametrize("cn_req", [0, 24])
@pytest.mark.parametrize("name,lmax,make", SHAPES)
def test_direction_cell_tables_are_proven_bounds(harness, name, lmax, make, cn_req):
    if cn_req and name != "perturbed_l30":
        pytest.skip("coarse cube map checked on one shape")
    a, b = make()
    nt, nphi = 16, 32
    err = C.create_string_buffer(256)
    t = harness.sth_build(lmax, a.ctypes.data, b.ctypes.data, 1.0, nt, nphi, cn_req, err, 256)
    assert t, err.value
    cn = harness.sth_cube_n(t)
    sc = np.zeros(13)
    harness.sth_scalars(t, sc.ctypes.data)
    rmax, rmin, h1, h2, step, pad, r_sup, r_inf = sc[0], sc[1], sc[3], sc[4], sc[5], sc[6], sc[7], sc[8]
    deltas = sc[9:13]
    nc = 6 * cn * cn
    ub2, lb2, wide2 = np.zeros(nc, np.float32), np.zeros(nc, np.float32), np.zeros(4 * nc, np.float32)
    harness.sth_cube(t, ub2.ctypes.data, lb2.ctypes.data, wide2.ctypes.data)
    o = O.Oracle()
    o.set_quadrature(nt, nphi)
    sid = o.add_shape(lmax, a, b, 1.0)
    # node tables and bounding radii bit-identical to the oracle's
    po = o.shape_props(sid)
    assert po["rmax"] == rmax and po["rmin"] == rmin
    pts, nds = np.zeros((nt * nphi, 3)), np.zeros((nt * nphi, 3))
    harness.sth_nodes(t, pts.ctypes.data, nds.ctypes.data)
    op, on = o.nodes(sid, nt * nphi)
    assert np.array_equal(pts, op) and np.array_equal(nds, on)

    # (1) derivative bounds recomputed from the coefficients
    H1, H2 = h_bounds(lmax, np.asarray(a), np.asarray(b))
    assert abs(h1 - H1) <= 1e-9 * H1 and abs(h2 - H2) <= 1e-9 * H2

    # (2) the bounds hold against numerical derivatives along random great circles (t = arc length)
    rng = np.random.default_rng(11)
    tt = np.linspace(0, 2 * np.pi, 4001)
    dt = tt[1] - tt[0]
    worst1 = worst2 = 0.0
    for _ in range(12):
        e1 = rng.normal(size=3); e1 /= np.linalg.norm(e1)
        e2 = np.cross(e1, rng.normal(size=3)); e2 /= np.linalg.norm(e2)
        r = o.sh_radius(sid, np.cos(tt)[:, None] * e1 + np.sin(tt)[:, None] * e2)
        worst1 = max(worst1, np.abs(np.gradient(r, dt)).max())
        worst2 = max(worst2, np.abs(r[2:] - 2 * r[1:-1] + r[:-2]).max() / dt ** 2)
    assert worst1 <= h1 and worst2 <= h2 * (1 + 1e-6) + 1e-6, (worst1, h1, worst2, h2)

    # (3) every table entry equals the formula, with the samples taken from the oracle
    sub = int(round(2.0 / step / cn))
    assert abs(step - 2.0 / (cn * sub)) < 1e-15
    assert abs(pad - (0.25 * h2 * step * step + h1 * 1e-5)) <= 1e-12
    g = -1.0 + np.arange(cn * sub + 1) * step
    uu, vv = np.meshgrid(g, g, indexing="ij")
    for f in range(6):
        rs = o.sh_radius(sid, face_dirs(f, uu, vv).reshape(-1, 3)).reshape(uu.shape)
        win = np.lib.stride_tricks.sliding_window_view(rs, (sub + 1, sub + 1))[::sub, ::sub]
        mx, mn = win.max(axis=(2, 3)), win.min(axis=(2, 3))
        ub = np.minimum(mx + pad, rmax)
        lb = np.maximum(mn - pad, rmin)
        tu = np.sqrt(ub2[f * cn * cn:(f + 1) * cn * cn].astype(np.float64)).reshape(cn, cn)
        tl = np.sqrt(lb2[f * cn * cn:(f + 1) * cn * cn].astype(np.float64)).reshape(cn, cn)
        assert np.all(tu >= ub * (1 - 1e-9)) and np.all(tu <= ub * (1 + 2e-7) + 1e-9), (f, np.abs(tu - ub).max())
        assert np.all(tl <= lb * (1 + 1e-9)) and np.all(tl >= lb * (1 - 2e-7) - 1e-9), (f, np.abs(tl - lb).max())
    assert r_sup <= rmax and r_inf >= rmin          # rmax / rmin are proven, not sampled, bounds

    # (4) sampling: random directions and directions ON cell / face borders, FP32 cell index as in the kernels
    d = rng.normal(size=(300000, 3))
    gb = np.linspace(-1, 1, cn + 1)
    bu, bv = np.meshgrid(gb, np.linspace(-1, 1, 97))
    border = np.concatenate([np.stack([np.ones(bu.size), bu.ravel(), bv.ravel()], 1), np.stack([bu.ravel(), -np.ones(bu.size), bv.ravel()], 1),
                             np.stack([bv.ravel(), bu.ravel(), np.ones(bu.size)], 1)])
    d = np.concatenate([d, border, border * (1 + 1e-7 * rng.normal(size=border.shape))])
    d /= np.linalg.norm(d, axis=1)[:, None]
    scale = rng.uniform(0.3, 1.5, size=(len(d), 1))            # the cell depends on the direction only
    r = o.sh_radius(sid, d)
    cell = cube_cell(d * scale, cn)
    assert np.all(r * r <= ub2[cell].astype(np.float64)), float((r * r - ub2[cell]).max())
    assert np.all(r * r >= lb2[cell].astype(np.float64)), float((lb2[cell] - r * r).max())
    assert np.all((r >= rmin) & (r <= rmax))
    shell = np.sqrt(ub2[cell].astype(np.float64)) - np.sqrt(lb2[cell].astype(np.float64))
    assert np.median(shell) < 0.06 * rmax, np.median(shell)     # and not uselessly loose
    # candidate-cache tables: a node cached at direction d0 may drift by <= gamma before the cache is rebuilt
    for lv in range(4):
        delta = deltas[lv]
        gamma = np.arcsin(min(1.0, delta / (rmin + 2 * delta)))
        d0 = d[:60000]
        axis = np.cross(d0, rng.normal(size=d0.shape)); axis /= np.linalg.norm(axis, axis=1)[:, None]
        ang = rng.uniform(0, gamma, size=(len(d0), 1))
        d1 = d0 * np.cos(ang) + np.cross(axis, d0) * np.sin(ang)       # rotated by <= gamma
        need = (o.sh_radius(sid, d1) + delta) ** 2
        wide = wide2[lv * nc:(lv + 1) * nc]
        assert np.all(need <= wide[cube_cell(d0, cn)].astype(np.float64)), float((need / wide[cube_cell(d0, cn)]).max())
        assert np.all(wide >= ub2)
    harness.sth_free(t)
    o.close()


def test_rough_shape_is_refused_or_bounded(harness):
    """A shape whose bounding radius cannot be proven from the 0.5 % pad must fail loudly, not silently."""
    lmax = 30
    a, b = W.sphere_shape(lmax, 1.0)
    a = np.array(a); b = np.array(b)
    a[lmax * (lmax + 1) // 2 + 7] = 0.2          # strong l = 30 ripple
    err = C.create_string_buffer(256)
    t = harness.sth_build(lmax, a.ctypes.data, b.ctypes.data, 1.0, 16, 32, 0, err, 256)
    if t:
        harness.sth_free(t)
    else:
        assert b"cannot prove" in err.value
