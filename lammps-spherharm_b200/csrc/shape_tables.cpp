// shape_tables.cpp — host-side per-shape tables.
// Reference source for `atom_style spherharm`: NOT IN MOUNT (see include/shgpu.h header note).
// Compiled with -ffp-contract=off; no fused operations on this path.
#include "shape_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace shgpu {

static const double kPi = 3.14159265358979323846;

// Legendre polynomial P_n(z) and the pair (P_n, P_{n-1}) by upward recurrence.
static inline void legendre_pn(int n, double z, double &pn, double &pnm1) {
  double p1 = 1.0, p2 = 0.0;
  for (int j = 1; j <= n; j++) {
    double p3 = p2;
    p2 = p1;
    p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
  }
  pn = p1;
  pnm1 = p2;
}

void gauss_legendre_nodes(int n, std::vector<double> &x, std::vector<double> &w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  for (int i = 0; i < (n + 1) / 2; i++) {
    double z = std::cos(kPi * (i + 0.75) / (n + 0.5));
    double pn, pnm1, deriv = 1.0;
    for (int it = 0; it < 100; it++) {
      legendre_pn(n, z, pn, pnm1);
      deriv = n * (z * pn - pnm1) / (z * z - 1.0);
      double zprev = z;
      z = zprev - pn / deriv;
      if (std::fabs(z - zprev) < 1e-15) break;
    }
    legendre_pn(n, z, pn, pnm1);
    deriv = n * (z * pn - pnm1) / (z * z - 1.0);
    x[i] = -z;
    x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * deriv * deriv);
  }
}

static inline int lm(int l, int m) { return l * (l + 1) / 2 + m; }

void legendre_normalised(int lmax, double x, std::vector<double> &P) {
  P.resize((size_t)(lmax + 1) * (lmax + 2) / 2);
  const double s = std::sqrt((1.0 - x) * (1.0 + x));
  double sect = std::sqrt(1.0 / (4.0 * kPi));
  for (int m = 0; m <= lmax; m++) {
    if (m > 0) sect = std::sqrt((2.0 * m + 1.0) / (2.0 * m)) * s * sect;
    P[lm(m, m)] = sect;
    if (m < lmax) P[lm(m + 1, m)] = std::sqrt(2.0 * m + 3.0) * x * sect;
    for (int l = m + 2; l <= lmax; l++) {
      const double A = std::sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
      const double B = std::sqrt((((double)l - 1.0) * (l - 1.0) - (double)m * m) /
                                 (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
      P[lm(l, m)] = A * (x * P[lm(l - 1, m)] - B * P[lm(l - 2, m)]);
    }
  }
}

void rotation_from_quat(const double q[4], double R[3][3]) {
  const double w2 = q[0] * q[0], i2 = q[1] * q[1], j2 = q[2] * q[2], k2 = q[3] * q[3];
  const double twoij = (2.0 * q[1]) * q[2], twoik = (2.0 * q[1]) * q[3], twojk = (2.0 * q[2]) * q[3];
  const double twoiw = (2.0 * q[1]) * q[0], twojw = (2.0 * q[2]) * q[0], twokw = (2.0 * q[3]) * q[0];
  R[0][0] = ((w2 + i2) - j2) - k2;
  R[0][1] = twoij - twokw;
  R[0][2] = twojw + twoik;
  R[1][0] = twoij + twokw;
  R[1][1] = ((w2 - i2) + j2) - k2;
  R[1][2] = twojk - twoiw;
  R[2][0] = twoik - twojw;
  R[2][1] = twojk + twoiw;
  R[2][2] = ((w2 - i2) - j2) + k2;
}

namespace {

struct RadiusJet { double r, r_theta, r_phi; };

// setup-path series for r and its angular derivatives from the raw coefficients
RadiusJet series_with_derivatives(const ShapeTables &s, std::vector<double> &P, double theta, double phi) {
  const int L = s.lmax;
  const double x = std::cos(theta), st = std::sin(theta);
  legendre_normalised(L, x, P);
  double rr = 0, rt = 0, rp = 0;
  for (int l = 0; l <= L; l++)
    for (int m = 0; m <= l; m++) {
      const int k = lm(l, m);
      const double cm = std::cos(m * phi), sm = std::sin(m * phi);
      const double Pl = P[k];
      const double Pl1 = (l > m) ? P[lm(l - 1, m)] : 0.0;
      const double flm =
          (l > m) ? std::sqrt((2.0 * l + 1.0) * ((double)l * l - (double)m * m) / (2.0 * l - 1.0)) : 0.0;
      const double dP = (l * x * Pl - flm * Pl1) / st;
      const double ang = s.a_raw[k] * cm + s.b_raw[k] * sm;
      rr += Pl * ang;
      rt += dP * ang;
      rp += Pl * m * (s.b_raw[k] * cm - s.a_raw[k] * sm);
    }
  return {rr, rt, rp};
}

void fold_recurrence(ShapeTables &s) {
  const int L = s.lmax, T = s.nterms;
  s.Ap.assign(T, 0.0);
  s.ah.assign(T, 0.0);
  s.bh.assign(T, 0.0);
  std::vector<double> alpha(L + 1, 1.0);
  int o = 0;
  double cm = std::sqrt(1.0 / (4.0 * kPi));
  for (int m = 0; m <= L; m++) {
    if (m > 0) cm = cm * std::sqrt((2.0 * m + 1.0) / (2.0 * m));
    for (int l = m; l <= L; l++) {
      const int idx = o + (l - m), k = lm(l, m);
      double Ap;
      if (l == m) {
        alpha[l] = 1.0;
        Ap = 0.0;
      } else if (l == m + 1) {
        alpha[l] = 1.0;
        Ap = std::sqrt(2.0 * m + 3.0);
      } else {
        const double A = std::sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
        const double B = std::sqrt((((double)l - 1.0) * (l - 1.0) - (double)m * m) /
                                   (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
        alpha[l] = (A * B) * alpha[l - 2];
        Ap = (A * alpha[l - 1]) / alpha[l];
      }
      s.Ap[idx] = Ap;
      s.ah[idx] = (s.a_raw[k] * alpha[l]) * cm;
      s.bh[idx] = (m == 0) ? 0.0 : (s.b_raw[k] * alpha[l]) * cm;
    }
    o += L + 1 - m;
  }
}

// r(direction) with the folded tables; plain a*b+c arithmetic (bounds only, not the decision path)
double radius_for_bounds(const ShapeTables &s, double s0, double s1, double s2) {
  const int L = s.lmax;
  const double rho = std::sqrt(s0 * s0 + s1 * s1 + s2 * s2), inv = 1.0 / rho;
  const double x = s2 * inv, zx = s0 * inv, zy = s1 * inv;
  double u = 1.0, v = 0.0, r = 0.0;
  int base = 0;
  for (int m = 0; m <= L; m++) {
    if (m > 0) { const double un = u * zx - v * zy, vn = u * zy + v * zx; u = un; v = vn; }
    double C = s.ah[base], S = s.bh[base];
    if (m < L) {
      double q1 = s.Ap[base + 1] * x, q2 = 1.0;
      C += s.ah[base + 1] * q1; S += s.bh[base + 1] * q1;
      for (int i = 2; i <= L - m; i++) {
        const double q = s.Ap[base + i] * x * q1 - q2;
        C += s.ah[base + i] * q; S += s.bh[base + i] * q;
        q2 = q1; q1 = q;
      }
    }
    r += u * C + v * S;
    base += L + 1 - m;
  }
  return r;
}

// cube-map direction cells: face f in 0..5 = +x,-x,+y,-y,+z,-z; (u,v) = the two other components (in x,y,z
// order) divided by |major|; cell (iu,iv) = floor((u+1)/2*n).  Samples overlap one sub-step into the
// neighbouring cells / faces and the bound is padded by the largest adjacent-sample difference, so that an
// FP32 cell index (pair_warp_kernel.cuh) that lands in a neighbouring cell near a border is still covered.
void build_cube_bounds(ShapeTables &s, int n, int sub) {
  s.cube_n = n;
  s.cube_bound2.assign((size_t)6 * n * n, 0.0f);
  const int ns = n * sub + 3;                  // samples per face edge: indices -1 .. n*sub+1
  std::vector<double> r((size_t)ns * ns);
  for (int f = 0; f < 6; f++) {
    const int major = f / 2;
    const double sgn = (f % 2) ? -1.0 : 1.0;
    for (int a = 0; a < ns; a++)
      for (int b = 0; b < ns; b++) {
        const double u = -1.0 + 2.0 * (a - 1) / (double)(n * sub), v = -1.0 + 2.0 * (b - 1) / (double)(n * sub);
        double d[3];
        d[major] = sgn;
        d[major == 0 ? 1 : 0] = u;
        d[major == 2 ? 1 : 2] = v;
        r[(size_t)a * ns + b] = radius_for_bounds(s, d[0], d[1], d[2]);
      }
    for (int iu = 0; iu < n; iu++)
      for (int iv = 0; iv < n; iv++) {
        double mx = 0, dmax = 0;
        for (int a = iu * sub; a <= (iu + 1) * sub + 2; a++)
          for (int b = iv * sub; b <= (iv + 1) * sub + 2; b++) {
            const double v0 = r[(size_t)a * ns + b];
            mx = std::max(mx, v0);
            if (a + 1 < ns) dmax = std::max(dmax, std::fabs(r[(size_t)(a + 1) * ns + b] - v0));
            if (b + 1 < ns) dmax = std::max(dmax, std::fabs(r[(size_t)a * ns + b + 1] - v0));
          }
        const double bound = (mx + dmax) * (1.0 + 1e-6);
        float b2 = (float)(bound * bound);
        b2 = std::nextafter(b2, 3.0e38f);      // round up: the FP32 value must not be below the FP64 bound^2
        s.cube_bound2[((size_t)f * n + iu) * n + iv] = b2;
      }
  }
}

// Table for the candidate cache (pair_split_kernels.cuh).  A node cached at build time with direction d0 (as
// seen from this shape's centre) and radius rho0 may move by at most delta before the cache is rebuilt, so its
// direction drifts by at most gamma = asin(delta / (rmin + 2 delta)) (nodes closer than rmin + 2 delta are cached
// unconditionally).  wide2(cell) = (sqrt(max cube_bound2 over all cells within gamma of the cell) + delta)^2.
void build_cache_table(ShapeTables &s) {
  const int n = s.cube_n;
  s.cache_delta = 0.02 * s.rmax;
  const double ratio = s.cache_delta / (s.rmin + 2.0 * s.cache_delta);
  const double gamma = 1.1 * std::asin(std::min(1.0, ratio)) + 1e-3;
  std::vector<double> dir((size_t)6 * n * n * 3), hd((size_t)6 * n * n);
  auto cell_dir = [&](int f, double u, double v, double *d) {
    const int major = f / 2;
    d[major] = (f % 2) ? -1.0 : 1.0;
    d[major == 0 ? 1 : 0] = u;
    d[major == 2 ? 1 : 2] = v;
    const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int k = 0; k < 3; k++) d[k] /= nn;
  };
  const double du = 2.0 / n, ov = du / 6.0;   // cells cover their own extent plus the one-sub-step overlap
  for (int f = 0; f < 6; f++)
    for (int iu = 0; iu < n; iu++)
      for (int iv = 0; iv < n; iv++) {
        const size_t c = ((size_t)f * n + iu) * n + iv;
        const double u0 = -1.0 + iu * du, v0 = -1.0 + iv * du;
        cell_dir(f, u0 + 0.5 * du, v0 + 0.5 * du, &dir[3 * c]);
        double worst = 0;
        for (int a = 0; a < 2; a++)
          for (int b = 0; b < 2; b++) {
            double cd[3];
            cell_dir(f, u0 - ov + a * (du + 2 * ov), v0 - ov + b * (du + 2 * ov), cd);
            const double dot = cd[0] * dir[3 * c] + cd[1] * dir[3 * c + 1] + cd[2] * dir[3 * c + 2];
            worst = std::max(worst, std::acos(std::min(1.0, std::max(-1.0, dot))));
          }
        hd[c] = worst;   // angular half-diagonal of the (overlapped) cell
      }
  const size_t nc = (size_t)6 * n * n;
  s.cube_wide2.assign(nc, 0.0f);
  double hdmax = 0;
  for (size_t c = 0; c < nc; c++) hdmax = std::max(hdmax, hd[c]);
  const double reach = gamma + 2.0 * hdmax + 1e-6;
  const double cos_reach = reach >= 3.14159 ? -2.0 : std::cos(reach);
  for (size_t c = 0; c < nc; c++) {
    float mx = 0.0f;
    for (size_t e = 0; e < nc; e++) {
      const double dot = dir[3 * c] * dir[3 * e] + dir[3 * c + 1] * dir[3 * e + 1] + dir[3 * c + 2] * dir[3 * e + 2];
      if (dot >= cos_reach) mx = std::max(mx, s.cube_bound2[e]);
    }
    const double w = std::sqrt((double)mx) + s.cache_delta;
    float w2 = (float)(w * w * (1.0 + 1e-6));
    s.cube_wide2[c] = std::nextafter(w2, 3.0e38f);
  }
}

// cyclic Jacobi for a symmetric 3x3 (principal inertia axes)
void jacobi_sym3(double A[3][3], double ev[3], double V[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 50; sweep++) {
    const double offd = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
    const double diag = std::fabs(A[0][0]) + std::fabs(A[1][1]) + std::fabs(A[2][2]);
    if (offd <= 1e-15 * diag) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (std::fabs(A[p][q]) <= 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 3; k++) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = cs * akp - sn * akq;
          A[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = cs * apk - sn * aqk;
          A[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = cs * vkp - sn * vkq;
          V[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int i = 0; i < 3; i++) ev[i] = A[i][i];
}

void quat_from_rotation(double R[3][3], double q[4]) {
  const double tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0) {
    const double s = std::sqrt(tr + 1.0) * 2;
    q[0] = 0.25 * s; q[1] = (R[2][1] - R[1][2]) / s; q[2] = (R[0][2] - R[2][0]) / s; q[3] = (R[1][0] - R[0][1]) / s;
  } else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) {
    const double s = std::sqrt(1.0 + R[0][0] - R[1][1] - R[2][2]) * 2;
    q[0] = (R[2][1] - R[1][2]) / s; q[1] = 0.25 * s; q[2] = (R[0][1] + R[1][0]) / s; q[3] = (R[0][2] + R[2][0]) / s;
  } else if (R[1][1] > R[2][2]) {
    const double s = std::sqrt(1.0 + R[1][1] - R[0][0] - R[2][2]) * 2;
    q[0] = (R[0][2] - R[2][0]) / s; q[1] = (R[0][1] + R[1][0]) / s; q[2] = 0.25 * s; q[3] = (R[1][2] + R[2][1]) / s;
  } else {
    const double s = std::sqrt(1.0 + R[2][2] - R[0][0] - R[1][1]) * 2;
    q[0] = (R[1][0] - R[0][1]) / s; q[1] = (R[0][2] + R[2][0]) / s; q[2] = (R[1][2] + R[2][1]) / s; q[3] = 0.25 * s;
  }
  const double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] /= nrm;
}

}  // namespace

std::string build_shape_tables(int lmax, const double *a_lm, const double *b_lm, double density, int n_theta,
                               int n_phi, ShapeTables &s) {
  if (lmax < 0 || lmax > 128) return "lmax out of range";
  if (!(density > 0)) return "density must be > 0";
  if (!a_lm) return "a_lm is NULL";
  if ((long)n_theta * n_phi > 65535) return "quadrature has more than 65535 nodes";
  const int L = lmax, T = (L + 1) * (L + 2) / 2;
  s = ShapeTables();
  s.lmax = L; s.nterms = T; s.density = density;
  s.n_theta = n_theta; s.n_phi = n_phi; s.nq = n_theta * n_phi;
  s.a_raw.assign(a_lm, a_lm + T);
  if (b_lm) s.b_raw.assign(b_lm, b_lm + T); else s.b_raw.assign(T, 0.0);
  for (int l = 0; l <= L; l++) s.b_raw[lm(l, 0)] = 0.0;
  fold_recurrence(s);

  const int nt = n_theta, np = n_phi, nq = s.nq;
  for (int d = 0; d < 3; d++) { s.node_p[d].assign(nq, 0.0); s.node_n[d].assign(nq, 0.0); }
  std::vector<double> gx, gw, P;
  gauss_legendre_nodes(nt, gx, gw);
  s.row_x = gx;
  const double dphi = 2.0 * kPi / np;
  double vol = 0, m1[3] = {0, 0, 0}, Io[3][3] = {{0}};
  double rmax = 0, rmin = 1e300;
  for (int a = 0; a < nt; a++) {
    const double theta = std::acos(gx[a]);
    for (int b = 0; b < np; b++) {
      const double phi = (b + 0.5) * dphi;
      const RadiusJet j = series_with_derivatives(s, P, theta, phi);
      const double r = j.r, rth = j.r_theta, rph = j.r_phi;
      if (!(r > 0)) return "shape not star-shaped: r <= 0 at a node";
      const double st = std::sin(theta), ct = std::cos(theta), cp = std::cos(phi), sp = std::sin(phi);
      const double rh[3] = {st * cp, st * sp, ct}, th[3] = {ct * cp, ct * sp, -st}, ph[3] = {-sp, cp, 0.0};
      const double w = gw[a] * dphi;
      const int k = a * np + b;
      for (int d = 0; d < 3; d++) {
        s.node_p[d][k] = r * rh[d];
        s.node_n[d][k] = (r * r * rh[d] - r * rth * th[d] - (r * rph / st) * ph[d]) * w;
      }
      const double r3 = r * r * r;
      vol += w * r3 / 3.0;
      for (int d = 0; d < 3; d++) m1[d] += w * (r3 * r / 4.0) * rh[d];
      const double r5 = r3 * r * r / 5.0;
      for (int d = 0; d < 3; d++)
        for (int e = 0; e < 3; e++) Io[d][e] += w * r5 * ((d == e) - rh[d] * rh[e]);
      if (r > rmax) rmax = r;
      if (r < rmin) rmin = r;
    }
  }
  const int dt_ = 4 * nt, dp_ = 4 * np;
  for (int a = 0; a < dt_; a++)
    for (int b = 0; b < dp_; b++) {
      const RadiusJet j = series_with_derivatives(s, P, (a + 0.5) * kPi / dt_, (b + 0.5) * 2.0 * kPi / dp_);
      if (!(j.r > 0)) return "shape not star-shaped: r <= 0";
      if (j.r > rmax) rmax = j.r;
      if (j.r < rmin) rmin = j.r;
    }
  s.rmax = 1.005 * rmax;
  s.rmin = 0.995 * rmin;
  s.volume = vol;
  s.mass = density * vol;
  for (int d = 0; d < 3; d++) s.com[d] = m1[d] / vol;
  double Ic[3][3];
  const double c2 = s.com[0] * s.com[0] + s.com[1] * s.com[1] + s.com[2] * s.com[2];
  for (int d = 0; d < 3; d++)
    for (int e = 0; e < 3; e++) Ic[d][e] = density * Io[d][e] - s.mass * (c2 * (d == e) - s.com[d] * s.com[e]);
  for (int d = 0; d < 3; d++)
    for (int e = d + 1; e < 3; e++) {
      const double av = 0.5 * (Ic[d][e] + Ic[e][d]);
      Ic[d][e] = Ic[e][d] = av;
    }
  double ev[3], V[3][3];
  jacobi_sym3(Ic, ev, V);
  V[0][2] = V[1][0] * V[2][1] - V[2][0] * V[1][1];
  V[1][2] = V[2][0] * V[0][1] - V[0][0] * V[2][1];
  V[2][2] = V[0][0] * V[1][1] - V[1][0] * V[0][1];
  for (int d = 0; d < 3; d++) s.inertia[d] = ev[d];
  double qp[4];
  quat_from_rotation(V, qp);
  for (int d = 0; d < 4; d++) s.quat_principal[d] = qp[d];
  rotation_from_quat(qp, s.Rp);
  build_cube_bounds(s, 24, 6);
  build_cache_table(s);
  return "";
}

}  // namespace shgpu
