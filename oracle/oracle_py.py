"""ctypes binding of the CPU oracle (oracle/libshoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED (see sh_oracle.h): the reference mount has no source to check against.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_int64)


def build(force=False):
    so = os.path.join(_HERE, "libshoracle.so")
    src = os.path.join(_HERE, "sh_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_last_error.restype = C.c_char_p
        _LIB.orc_last_error.argtypes = [C.c_void_p]
    return _LIB


def _d(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=c_dp):
    return None if a is None else a.ctypes.data_as(t)


class Oracle:
    """Same call surface as lammps_spherharm_b200.capi.ShGpu (sh_* C-ABI), CPU FP64."""

    def __init__(self, threads=1):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.L.orc_set_threads(self.h, int(threads))
        self.n = 0

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error(self.h).decode())

    def set_box(self, lo, hi, periodic):
        lo, hi = _d(lo), _d(hi)
        per = np.ascontiguousarray(periodic, dtype=np.int32)
        self._ck(self.L.orc_set_box(self.h, _p(lo), _p(hi), _p(per, c_ip)))

    def set_quadrature(self, nt, nphi):
        self._ck(self.L.orc_set_quadrature(self.h, int(nt), int(nphi)))

    def add_shape(self, lmax, a_lm, b_lm, density=1.0):
        a, b = _d(a_lm), _d(b_lm)
        sid = C.c_int(-1)
        self._ck(self.L.orc_add_shape(self.h, int(lmax), _p(a), _p(b), C.c_double(density), C.byref(sid)))
        return sid.value

    def shape_props(self, shape):
        vol, rmax, rmin = C.c_double(), C.c_double(), C.c_double()
        com, inertia, qp = np.zeros(3), np.zeros(3), np.zeros(4)
        self._ck(self.L.orc_get_shape_props(self.h, int(shape), C.byref(vol), _p(com), _p(inertia), _p(qp),
                                            C.byref(rmax), C.byref(rmin)))
        return dict(volume=vol.value, com=com, inertia=inertia, quat_principal=qp, rmax=rmax.value, rmin=rmin.value)

    def nodes(self, shape, nq):
        p, nds = np.zeros((nq, 3)), np.zeros((nq, 3))
        self._ck(self.L.orc_get_nodes(self.h, int(shape), _p(p), _p(nds)))
        return p, nds

    def set_atoms(self, shape, x, v=None, quat=None, angmom=None, tag=None):
        shape = np.ascontiguousarray(shape, dtype=np.int32)
        n = len(shape)
        x, v, quat, angmom = _d(x), _d(v), _d(quat), _d(angmom)
        tag = None if tag is None else np.ascontiguousarray(tag, dtype=np.int64)
        self._ck(self.L.orc_set_atoms(self.h, C.c_int64(n), _p(tag, c_lp), _p(shape, c_ip), _p(x), _p(v), _p(quat),
                                      _p(angmom)))
        self.n = n

    def pair_coeff(self, si, sj, k, exponent):
        self._ck(self.L.orc_pair_coeff(self.h, int(si), int(sj), C.c_double(k), C.c_double(exponent)))

    def add_wall(self, point, normal, k, exponent):
        p, nn = _d(point), _d(normal)
        self._ck(self.L.orc_add_wall(self.h, _p(p), _p(nn), C.c_double(k), C.c_double(exponent)))

    def set_gravity(self, g):
        g = _d(g)
        self._ck(self.L.orc_set_gravity(self.h, _p(g)))

    def set_neighbor(self, skin, every=1, check=1):
        self._ck(self.L.orc_set_neighbor(self.h, C.c_double(skin), int(every), int(check)))

    def set_damping(self, gamma_lin, gamma_rot):
        self._ck(self.L.orc_set_damping(self.h, C.c_double(gamma_lin), C.c_double(gamma_rot)))

    def pair_dissipation(self, si, sj, gamma_n, gamma_t, mu):
        self._ck(self.L.orc_pair_dissipation(self.h, int(si), int(sj), C.c_double(gamma_n), C.c_double(gamma_t), C.c_double(mu)))

    def get_stress(self):
        w, k = np.zeros(9), np.zeros(9)
        self._ck(self.L.orc_get_stress(self.h, _p(w), _p(k)))
        return dict(virial=w.reshape(3, 3), kinetic=k.reshape(3, 3))

    def set_shear(self, rate):
        self._ck(self.L.orc_set_shear(self.h, C.c_double(rate)))

    def set_timestep(self, dt):
        self._ck(self.L.orc_set_timestep(self.h, C.c_double(dt)))

    def compute_forces(self):
        self._ck(self.L.orc_compute_forces(self.h))

    def run(self, nsteps):
        self._ck(self.L.orc_run(self.h, C.c_int64(nsteps)))

    # ---- multi-rank test support: the surface decomp.py drives (host pointers instead of device pointers)
    def set_ghost_count(self, nghost):
        self._ck(self.L.orc_set_ghost_count(self.h, C.c_int64(nghost)))

    def step_begin(self):
        flag = C.c_int(0)
        self._ck(self.L.orc_step_begin(self.h, C.byref(flag)))
        return flag.value

    def step_end(self, rebuild):
        self._ck(self.L.orc_step_end(self.h, int(rebuild)))

    def synchronize(self):
        pass

    def pack_atoms(self, m, idx_ptr, shift_ptr, out_ptr):
        self._ck(self.L.orc_pack_atoms(self.h, C.c_int64(m), C.c_void_p(idx_ptr), C.c_void_p(shift_ptr), C.c_void_p(out_ptr)))

    def unpack_ghosts(self, first, m, in_ptr):
        self._ck(self.L.orc_unpack_ghosts(self.h, C.c_int64(first), C.c_int64(m), C.c_void_p(in_ptr)))

    def get_atoms(self, fields=None):
        n = self.n
        out = dict(x=np.zeros((n, 3)), v=np.zeros((n, 3)), quat=np.zeros((n, 4)), angmom=np.zeros((n, 3)),
                   f=np.zeros((n, 3)), torque=np.zeros((n, 3)))
        self._ck(self.L.orc_get_atoms(self.h, C.c_int64(n), _p(out["x"]), _p(out["v"]), _p(out["quat"]),
                                      _p(out["angmom"]), _p(out["f"]), _p(out["torque"])))
        return out

    def get_pairs(self):
        npairs = C.c_int64(0)
        self._ck(self.L.orc_get_pairs(self.h, C.c_int64(0), C.byref(npairs), None, None, None, None, None, None, None))
        m = npairs.value
        ti, tj = np.zeros(m, dtype=np.int64), np.zeros(m, dtype=np.int64)
        V, F, tau_i, tau_j, xc = np.zeros(m), np.zeros((m, 3)), np.zeros((m, 3)), np.zeros((m, 3)), np.zeros((m, 3))
        self._ck(self.L.orc_get_pairs(self.h, C.c_int64(m), C.byref(npairs), _p(ti, c_lp), _p(tj, c_lp), _p(V), _p(F),
                                      _p(tau_i), _p(tau_j), _p(xc)))
        return dict(tag_i=ti, tag_j=tj, V=V, F=F, tau_i=tau_i, tau_j=tau_j, centroid=xc)

    def get_counters(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.L.orc_get_counters(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(pair_evals=a.value, nodes_transformed=b.value, nodes_evaluated=c.value, nodes_inside=d.value)

    def get_energy(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.orc_get_energy(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(ke_trans=a.value, ke_rot=b.value, e_contact=c.value)

    def sh_radius(self, shape, dirs):
        dirs = _d(dirs)
        r = np.zeros(len(dirs))
        self._ck(self.L.orc_sh_radius(self.h, int(shape), C.c_int64(len(dirs)), _p(dirs), _p(r)))
        return r


def legendre_norm(lmax, x):
    P = np.zeros((lmax + 1) * (lmax + 2) // 2)
    lib().orc_legendre_norm(int(lmax), C.c_double(x), _p(P))
    return P


def gauss_legendre(n):
    x, w = np.zeros(n), np.zeros(n)
    lib().orc_gauss_legendre(int(n), _p(x), _p(w))
    return x, w


def project_ellipsoid(lmax, a, b, c, n_theta=128, n_phi=256):
    T = (lmax + 1) * (lmax + 2) // 2
    alm, blm = np.zeros(T), np.zeros(T)
    lib().orc_project_ellipsoid(int(lmax), C.c_double(a), C.c_double(b), C.c_double(c), int(n_theta), int(n_phi),
                                _p(alm), _p(blm))
    return alm, blm
