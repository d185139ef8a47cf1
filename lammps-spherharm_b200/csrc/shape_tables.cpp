// shape_tables.cpp — host-side per-shape tables.
// Reference source for `atom_style spherharm`: NOT IN MOUNT (see include/shgpu.h header note).
// Compiled with -ffp-contract=off; no fused operations on this path.
#include "shape_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

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

// ---- rigorous direction-cell tables ------------------------------------------------------------
// Derivative bounds.  Write r = sum_l f_l with f_l the degree-l part.  In the orthonormal real basis the
// coefficient vector of f_l is c_l = (a_l0, a_lm/sqrt2, b_lm/sqrt2), so by Cauchy-Schwarz and the addition theorem
// (sum_m Y_lm^2 = (2l+1)/4pi)   |f_l| <= B_l := sqrt((2l+1)/4pi) |c_l|_2   everywhere.  Restricted to a great circle
// parametrised by arc length t, f_l is a trigonometric polynomial of degree <= l, so Bernstein's inequality gives
// |f_l'| <= l B_l and |f_l''| <= l^2 B_l.  Hence  |dr/dt| <= H1 = sum l B_l,  |d2r/dt2| <= H2 = sum l^2 B_l.
void derivative_bounds(ShapeTables &s) {
  double h1 = 0, h2 = 0;
  for (int l = 1; l <= s.lmax; l++) {
    double n2 = s.a_raw[lm(l, 0)] * s.a_raw[lm(l, 0)];
    for (int m = 1; m <= l; m++) n2 += 0.5 * (s.a_raw[lm(l, m)] * s.a_raw[lm(l, m)] + s.b_raw[lm(l, m)] * s.b_raw[lm(l, m)]);
    const double B = std::sqrt((2.0 * l + 1.0) / (4.0 * kPi)) * std::sqrt(n2);
    h1 += l * B;
    h2 += (double)l * l * B;
  }
  s.h1_bound = h1 * (1.0 + 1e-12);
  s.h2_bound = h2 * (1.0 + 1e-12);
}

// Cube map: face f in 0..5 = +x,-x,+y,-y,+z,-z; (u,v) = the two other components (in x,y,z order) divided by |major|
// (gnomonic coordinates); cell (iu,iv) = floor((u+1)/2*n).  Straight lines in (u,v) are great-circle arcs and the
// map (u,v) -> direction is a contraction, so on a gnomonic grid of spacing D every direction x inside a grid square
// satisfies  min(corners) - H2 D^2/4 <= r(x) <= max(corners) + H2 D^2/4  (interpolation error of a function with
// |f''| <= H2 on a segment of length <= D, applied along v = const and then along the two u = const edges).
// The FP32 cell index of the kernels can be off by one cell for directions within ~1e-6 of a border; the term
// H1 * 1e-5 covers a border strip of width 1e-5 in (u,v).
// The sample grid of every face extends K steps beyond the face (directions (1,u,v) with |u| > 1 are ordinary
// directions) so that the candidate-cache tables can look across face edges.
struct FaceSamples {
  int ns = 0, K = 0, NS = 0;     // ns = NS + 2K + 1 samples per edge, u_a = -1 + (a - K) * step
  double step = 0;
  std::vector<double> r[6];
};

static void face_direction(int f, double u, double v, double d[3]) {
  const int major = f / 2;
  d[major] = (f % 2) ? -1.0 : 1.0;
  d[major == 0 ? 1 : 0] = u;
  d[major == 2 ? 1 : 2] = v;
}

void sample_faces(const ShapeTables &s, int NS, int K, FaceSamples &F) {
  F.NS = NS; F.K = K; F.ns = NS + 2 * K + 1; F.step = 2.0 / NS;
  const int ns = F.ns;
  for (int f = 0; f < 6; f++) F.r[f].assign((size_t)ns * ns, 0.0);
  const int nthreads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++)
    pool.emplace_back([&, t]() {
      for (int job = t; job < 6 * ns; job += nthreads) {
        const int f = job / ns, a = job % ns;
        const double u = -1.0 + (a - K) * F.step;
        for (int b = 0; b < ns; b++) {
          double d[3];
          face_direction(f, u, -1.0 + (b - K) * F.step, d);
          F.r[f][(size_t)a * ns + b] = radius_for_bounds(s, d[0], d[1], d[2]);
        }
      }
    });
  for (auto &th : pool) th.join();
}

static inline float sq_up(double v) { float q = (float)(v * v); return std::nextafter(q, 3.0e38f); }
static inline float sq_down(double v) { float q = (float)(v * v); return std::nextafter(q, 0.0f); }

const double kCacheFrac[SH_CACHE_LEVELS] = {0.005, 0.01, 0.02, 0.04};   // cache_delta[level] / rmax

std::string build_cube_tables(ShapeTables &s, int n) {
  derivative_bounds(s);
  const int sub = std::max(n <= 144 ? 2 : 1, (144 + n / 2) / n), NS = n * sub;   // sample spacing <= 2/144 whatever the cell size
  const double step = 2.0 / NS;
  for (int lv = 0; lv < SH_CACHE_LEVELS; lv++) s.cache_delta[lv] = kCacheFrac[lv] * s.rmax;
  // direction drift of a node that moves by <= delta at distance >= rmin + 2 delta from the centre
  double gamma[SH_CACHE_LEVELS];
  for (int lv = 0; lv < SH_CACHE_LEVELS; lv++)
    gamma[lv] = std::asin(std::min(1.0, s.cache_delta[lv] / (s.rmin + 2.0 * s.cache_delta[lv]))) * (1.0 + 1e-9) + 1e-9;
  const double eps_uv = 1e-5;
  // the sample grid must reach as far beyond a face as the largest drift box (a corner cell at the widest margin)
  int K = 3;
  {
    const double g = gamma[SH_CACHE_LEVELS - 1];
    auto stretch = [&](double rho) { const double m = 1.0 + rho + eps_uv; return 1.0 + 2.0 * m * m; };
    double rho = g * stretch(0.0);
    for (int it = 0; it < 200 && g * stretch(rho) > rho; it++) rho = g * stretch(rho) * 1.02;
    if (g * stretch(rho) > rho) return "candidate-cache table: margin too wide for this shape (rmin too small)";
    K = (int)std::ceil((rho * 1.05 + eps_uv) / step) + 3;
  }
  FaceSamples F;
  sample_faces(s, NS, K, F);
  const int ns = F.ns;
  s.cube_n = n;
  s.sample_step = step;
  s.sample_pad = 0.25 * s.h2_bound * step * step + s.h1_bound * eps_uv;
  const size_t nc = (size_t)6 * n * n;
  s.cube_ub2.assign(nc, 0.0f); s.cube_lb2.assign(nc, 0.0f);
  for (int lv = 0; lv < SH_CACHE_LEVELS; lv++) s.cube_wide2[lv].assign(nc, 0.0f);
  double rsup = 0, rinf = 1e300;
  std::vector<double> ubv(nc), lbv(nc);
  for (int f = 0; f < 6; f++)
    for (int iu = 0; iu < n; iu++)
      for (int iv = 0; iv < n; iv++) {
        const size_t c = ((size_t)f * n + iu) * n + iv;
        const int a0 = K + iu * sub, a1 = K + (iu + 1) * sub, b0 = K + iv * sub, b1 = K + (iv + 1) * sub;
        double mx = 0, mn = 1e300;
        for (int a = a0; a <= a1; a++)
          for (int b = b0; b <= b1; b++) { const double v = F.r[f][(size_t)a * ns + b]; mx = std::max(mx, v); mn = std::min(mn, v); }
        ubv[c] = (mx + s.sample_pad) * (1.0 + 1e-12);
        lbv[c] = (mn - s.sample_pad) * (1.0 - 1e-12);
        rsup = std::max(rsup, ubv[c]); rinf = std::min(rinf, lbv[c]);
        // candidate-cache tables: every direction within gamma of the (border-extended) cell.  An arc of length g
        // stays within (u,v)-distance rho of its start if g * (1 + R2(rho)) <= rho, R2 = max u^2+v^2 over the box + rho.
        const double ulo = -1.0 + iu * (2.0 / n), uhi = ulo + 2.0 / n, vlo = -1.0 + iv * (2.0 / n), vhi = vlo + 2.0 / n;
        for (int lv = 0; lv < SH_CACHE_LEVELS; lv++) {
          auto stretch = [&](double rho) {
            const double um = std::max(std::fabs(ulo), std::fabs(uhi)) + rho + eps_uv, vm = std::max(std::fabs(vlo), std::fabs(vhi)) + rho + eps_uv;
            return 1.0 + um * um + vm * vm;
          };
          double rho = gamma[lv] * stretch(0.0);
          for (int it = 0; it < 60 && gamma[lv] * stretch(rho) > rho; it++) rho = gamma[lv] * stretch(rho) * 1.02;
          if (gamma[lv] * stretch(rho) > rho) return "candidate-cache table: drift radius did not converge";
          const int e = (int)std::ceil((rho + eps_uv) / step);
          if (a0 - e < 0 || b0 - e < 0 || a1 + e >= ns || b1 + e >= ns) return "candidate-cache table: sample grid too small";
          double wm = 0;
          for (int a = a0 - e; a <= a1 + e; a++)
            for (int b = b0 - e; b <= b1 + e; b++) wm = std::max(wm, F.r[f][(size_t)a * ns + b]);
          const double w = (wm + s.sample_pad) * (1.0 + 1e-12) + s.cache_delta[lv];
          s.cube_wide2[lv][c] = sq_up(w);
        }
      }
  s.r_sup = rsup; s.r_inf = rinf;
  // rmax / rmin (shared bit for bit with the oracle: 4x oversampled node grid, 0.5 % pad) must be PROVEN bounds
  if (!(rsup <= s.rmax)) return "cannot prove the bounding radius: sup r <= " + std::to_string(rsup) + " > rmax = " + std::to_string(s.rmax) + " (shape too rough for the 0.5 % pad)";
  if (!(rinf >= s.rmin)) return "cannot prove the inscribed radius: inf r >= " + std::to_string(rinf) + " < rmin = " + std::to_string(s.rmin) + " (shape too rough for the 0.5 % pad)";
  for (size_t c = 0; c < nc; c++) {
    s.cube_ub2[c] = sq_up(std::min(ubv[c], s.rmax));
    s.cube_lb2[c] = sq_down(std::max(lbv[c], s.rmin));
  }
  return "";
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
                               int n_phi, ShapeTables &s, int cube_n) {
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
  return build_cube_tables(s, cube_n > 0 ? cube_n : 144);
}

}  // namespace shgpu
