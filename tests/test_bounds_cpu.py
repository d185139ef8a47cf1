"""CPU tests of the conservative tables behind the pair phase's work reduction: the product's host-side
table builder (csrc/shape_tables.cpp, compiled into a test harness) against the oracle's exact radius evaluation.

* cube_bound2[cell(d)] >= r(d)^2 for every direction d, with the cell index computed in FP32 exactly as the kernels
  do (so the overlap into neighbouring cells that absorbs FP32 index errors is exercised too);
* the candidate-cache table dominates (sqrt(narrow bound of any cell a direction can drift into) + delta)^2;
* node tables / bounding radii equal the oracle's bit for bit (also checked on the GPU through the C ABI)."""
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
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(out),
                           os.path.join(ROOT, "tests", "shape_tables_harness.cpp"),
                           os.path.join(ROOT, "lammps-spherharm_b200", "csrc", "shape_tables.cpp")])
    lib = C.CDLL(str(out))
    lib.sth_build.restype = C.c_void_p
    lib.sth_build.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_int]
    for f in (lib.sth_free, lib.sth_cube_n, lib.sth_scalars, lib.sth_cube, lib.sth_nodes):
        f.argtypes = [C.c_void_p] + [C.c_void_p] * (2 if f in (lib.sth_cube, lib.sth_nodes) else 1 if f is lib.sth_scalars else 0)
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


SHAPES = [("ellipsoid_l20", 20, lambda: W.ellipsoid_shape(20)), ("perturbed_l12", 12, lambda: W.perturbed_shape(12, 31)),
          ("perturbed_l30", 30, lambda: W.perturbed_shape(30, 33)), ("flat_l16", 16, lambda: W.project(16, W.ellipsoid_radius(1.0, 0.9, 0.45)))]


@pytest.mark.parametrize("name,lmax,make", SHAPES)
def test_direction_cell_tables_are_upper_bounds(harness, name, lmax, make):
    a, b = make()
    nt, nphi = 16, 32
    err = C.create_string_buffer(256)
    t = harness.sth_build(lmax, a.ctypes.data, b.ctypes.data, 1.0, nt, nphi, err, 256)
    assert t, err.value
    cn = harness.sth_cube_n(t)
    sc = np.zeros(4)
    harness.sth_scalars(t, sc.ctypes.data)
    rmax, rmin, delta = sc[0], sc[1], sc[2]
    narrow, wide = np.zeros(6 * cn * cn, np.float32), np.zeros(6 * cn * cn, np.float32)
    harness.sth_cube(t, narrow.ctypes.data, wide.ctypes.data)
    o = O.Oracle()
    o.set_quadrature(nt, nphi)
    sid = o.add_shape(lmax, a, b, 1.0)
    # tables and radii bit-identical to the oracle's
    po = o.shape_props(sid)
    assert po["rmax"] == rmax and po["rmin"] == rmin
    pts, nds = np.zeros((nt * nphi, 3)), np.zeros((nt * nphi, 3))
    harness.sth_nodes(t, pts.ctypes.data, nds.ctypes.data)
    op, on = o.nodes(sid, nt * nphi)
    assert np.array_equal(pts, op) and np.array_equal(nds, on)
    # --- narrow table: upper bound of r^2 for random directions and for directions ON cell / face borders
    rng = np.random.default_rng(7)
    d = rng.normal(size=(300000, 3))
    g = np.linspace(-1, 1, cn + 1)
    uu, vv = np.meshgrid(g, np.linspace(-1, 1, 97))
    border = np.concatenate([np.stack([np.ones(uu.size), uu.ravel(), vv.ravel()], 1), np.stack([uu.ravel(), -np.ones(uu.size), vv.ravel()], 1),
                             np.stack([vv.ravel(), uu.ravel(), np.ones(uu.size)], 1)])
    d = np.concatenate([d, border, border * (1 + 1e-7 * rng.normal(size=border.shape))])
    d /= np.linalg.norm(d, axis=1)[:, None]
    scale = rng.uniform(0.3, 1.5, size=(len(d), 1))            # the cell depends on the direction only
    r = o.sh_radius(sid, d)
    cell = cube_cell(d * scale, cn)
    assert np.all(r * r <= narrow[cell].astype(np.float64)), float((r * r - narrow[cell]).max())
    assert np.all(narrow <= np.float32((rmax * 1.02) ** 2))     # and not uselessly loose
    tight = np.sqrt(narrow[cell].astype(np.float64)) / r
    assert np.median(tight) < 1.06, np.median(tight)
    # --- cache table: a node cached at direction d0 may drift by <= gamma before the cache is rebuilt
    gamma = np.arcsin(min(1.0, delta / (rmin + 2 * delta)))
    d0 = d[:60000]
    axis = np.cross(d0, rng.normal(size=d0.shape)); axis /= np.linalg.norm(axis, axis=1)[:, None]
    ang = rng.uniform(0, gamma, size=(len(d0), 1))
    d1 = d0 * np.cos(ang) + np.cross(axis, d0) * np.sin(ang)       # rotated by <= gamma
    need = (np.sqrt(narrow[cube_cell(d1, cn)].astype(np.float64)) + delta) ** 2
    assert np.all(need <= wide[cube_cell(d0, cn)].astype(np.float64) * (1 + 1e-6)), float((need / wide[cube_cell(d0, cn)]).max())
    harness.sth_free(t)
    o.close()
