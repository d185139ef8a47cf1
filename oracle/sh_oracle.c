/* sh_oracle.c — CPU FP64 oracle for the SPHERHARM contact hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sh_oracle.h).  PARITY UNPINNED: the reference
 * mount /root/reference contains only README.md:1-3, so no function below can
 * cite a reference source line.  Each function instead cites the paragraph of
 * SURVEY.md Appendix A (the written-out form of BASELINE.json:5) it follows.
 *
 * Arithmetic contract: compiled with -ffp-contract=off; every
 * fused multiply-add on the node-decision path is an explicit fma() in the same
 * position as the CUDA kernel's fma(), so the inside/outside decision of every
 * surface node is bit-identical on CPU and GPU.  Sums over nodes are sequential
 * here and tree-shaped on the GPU (differences ~1e-15 relative).
 */
#include "sh_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_SHAPES 64
#define MAX_WALLS 16
#define PI 3.14159265358979323846

typedef struct {
  int lmax, T;
  double *a, *b;        /* raw, index l(l+1)/2+m */
  int *off;             /* m-major offsets, off[m] + (l-m) */
  double *Ap, *ah, *bh; /* folded recurrence / coefficients, m-major */
  double density, volume, mass, com[3], inertia[3], qp[4], Rp[3][3], rmax, rmin;
  int nq;
  double *p, *nds;      /* node table: point and oriented area element, body frame */
} shape_t;

typedef struct {
  int i, j;
  double V, E, F[3], ti[3], tj[3], xc[3];
  int ninside;
} pairres_t;

typedef struct { double c[3], n[3], k, m; } wall_t;

struct orc_ctx {
  char err[256];
  double lo[3], hi[3]; int periodic[3];
  int n_theta, n_phi;
  int nshape; shape_t shp[MAX_SHAPES];
  double pk[MAX_SHAPES][MAX_SHAPES], pm[MAX_SHAPES][MAX_SHAPES];
  double gn[MAX_SHAPES][MAX_SHAPES], gt[MAX_SHAPES][MAX_SHAPES], mu[MAX_SHAPES][MAX_SHAPES];   /* dissipation (A.5b) */
  int nwall; wall_t wall[MAX_WALLS];
  double g[3], skin, dt, gamma_lin, gamma_rot; int nthreads;
  int64_t n; int64_t *tag; int *shape;
  double *x, *v, *q, *L, *f, *tq;   /* n*3, n*3, n*4, n*3, n*3, n*3 */
  double *Rs, *c;                   /* pose: n*9, n*3 */
  int64_t nghost; double *c0;        /* multi-rank: last nghost atoms are ghosts; SH origins at the last set_atoms */
  int64_t npair, cappair; pairres_t *pr;
  int64_t cnt_pairs, cnt_trans, cnt_eval, cnt_inside;
  double e_contact; int forces_valid;
  double shear_rate, time;           /* Lees-Edwards shear (flow x, gradient y): image offset = shear_rate * Ly * time */
};

static int fail(orc_ctx *c, const char *msg) { snprintf(c->err, sizeof c->err, "%s", msg); return -1; }

orc_ctx *orc_create(void) {
  orc_ctx *c = (orc_ctx *)calloc(1, sizeof *c);
  for (int k = 0; k < 3; k++) { c->lo[k] = -1e30; c->hi[k] = 1e30; }
  c->n_theta = 32; c->n_phi = 64; c->skin = 0.0; c->dt = 1e-4;
  for (int i = 0; i < MAX_SHAPES; i++) for (int j = 0; j < MAX_SHAPES; j++) { c->pk[i][j] = 1.0; c->pm[i][j] = 1.0; c->gn[i][j] = c->gt[i][j] = c->mu[i][j] = 0.0; }
  c->nthreads = 1;
  return c;
}
static void free_shape(shape_t *s) {
  free(s->a); free(s->b); free(s->off); free(s->Ap); free(s->ah); free(s->bh); free(s->p); free(s->nds);
}
void orc_destroy(orc_ctx *c) {
  if (!c) return;
  for (int i = 0; i < c->nshape; i++) free_shape(&c->shp[i]);
  free(c->tag); free(c->shape); free(c->x); free(c->v); free(c->q); free(c->L); free(c->f); free(c->tq);
  free(c->Rs); free(c->c); free(c->c0); free(c->pr); free(c);
}
const char *orc_last_error(const orc_ctx *c) { return c->err; }

/* ---------------- Gauss-Legendre nodes ---------------- */
int orc_gauss_legendre(int n, double *x, double *w) {
  for (int i = 0; i < (n + 1) / 2; i++) {
    double z = cos(PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; it++) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double z1 = z; z = z1 - p1 / pp;
      if (fabs(z - z1) < 1e-15) break;
    }
    /* one more evaluation of pp at the converged root for the weight */
    { double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0); }
    x[i] = -z; x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  return 0;
}

/* ---------------- fully normalised associated Legendre ---------------- */
int orc_legendre_norm(int lmax, double x, double *P) {
  double s = sqrt((1.0 - x) * (1.0 + x));
  double pmm = sqrt(1.0 / (4.0 * PI));
  for (int m = 0; m <= lmax; m++) {
    if (m > 0) pmm = sqrt((2.0 * m + 1.0) / (2.0 * m)) * s * pmm;
    P[m * (m + 1) / 2 + m] = pmm;
    if (m < lmax) P[(m + 1) * (m + 2) / 2 + m] = sqrt(2.0 * m + 3.0) * x * pmm;
    for (int l = m + 2; l <= lmax; l++) {
      double A = sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
      double B = sqrt((((double)l - 1.0) * (l - 1.0) - (double)m * m) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
      P[l * (l + 1) / 2 + m] = A * (x * P[(l - 1) * l / 2 + m] - B * P[(l - 2) * (l - 1) / 2 + m]);
    }
  }
  return 0;
}

/* setup-path evaluation of r, dr/dtheta, dr/dphi from the raw coefficients (A.1, A.3) */
static void shape_eval_setup(const shape_t *s, double *P, double theta, double phi,
                             double *r, double *rth, double *rph) {
  int L = s->lmax;
  double x = cos(theta), st = sin(theta);
  orc_legendre_norm(L, x, P);
  double rr = 0, rt = 0, rp = 0;
  for (int l = 0; l <= L; l++)
    for (int m = 0; m <= l; m++) {
      int k = l * (l + 1) / 2 + m;
      double cm = cos(m * phi), sm = sin(m * phi);
      double Pl = P[k];
      double Pl1 = (l > m) ? P[(l - 1) * l / 2 + m] : 0.0;
      double flm = (l > m) ? sqrt((2.0 * l + 1.0) * ((double)l * l - (double)m * m) / (2.0 * l - 1.0)) : 0.0;
      double dP = (l * x * Pl - flm * Pl1) / st;
      double ang = s->a[k] * cm + s->b[k] * sm;
      rr += Pl * ang;
      rt += dP * ang;
      rp += Pl * m * (s->b[k] * cm - s->a[k] * sm);
    }
  *r = rr; *rth = rt; *rph = rp;
}

/* ---------------- small linear algebra ---------------- */
/* rotation matrix of unit quaternion (w,x,y,z); plain mul/add in this order */
static void quat_to_mat(const double q[4], double R[3][3]) {
  double w2 = q[0] * q[0], i2 = q[1] * q[1], j2 = q[2] * q[2], k2 = q[3] * q[3];
  double twoij = (2.0 * q[1]) * q[2], twoik = (2.0 * q[1]) * q[3], twojk = (2.0 * q[2]) * q[3];
  double twoiw = (2.0 * q[1]) * q[0], twojw = (2.0 * q[2]) * q[0], twokw = (2.0 * q[3]) * q[0];
  R[0][0] = ((w2 + i2) - j2) - k2; R[0][1] = twoij - twokw;         R[0][2] = twojw + twoik;
  R[1][0] = twoij + twokw;         R[1][1] = ((w2 - i2) + j2) - k2; R[1][2] = twojk - twoiw;
  R[2][0] = twoik - twojw;         R[2][1] = twojk + twoiw;         R[2][2] = ((w2 - i2) - j2) + k2;
}

static void jacobi3(double A[3][3], double ev[3], double V[3][3]) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 50; sweep++) {
    double offd = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (offd <= 1e-15 * diag) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (fabs(A[p][q]) <= 1e-300) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 3; k++) { double akp = A[k][p], akq = A[k][q]; A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq; }
        for (int k = 0; k < 3; k++) { double apk = A[p][k], aqk = A[q][k]; A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk; }
        for (int k = 0; k < 3; k++) { double vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
      }
  }
  for (int i = 0; i < 3; i++) ev[i] = A[i][i];
}

/* quaternion of a proper rotation matrix (columns = images of the basis vectors) */
static void mat_to_quat(double R[3][3], double q[4]) {
  double tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0) { double s = sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[2][1] - R[1][2]) / s; q[2] = (R[0][2] - R[2][0]) / s; q[3] = (R[1][0] - R[0][1]) / s; }
  else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) { double s = sqrt(1.0 + R[0][0] - R[1][1] - R[2][2]) * 2; q[0] = (R[2][1] - R[1][2]) / s; q[1] = 0.25 * s; q[2] = (R[0][1] + R[1][0]) / s; q[3] = (R[0][2] + R[2][0]) / s; }
  else if (R[1][1] > R[2][2]) { double s = sqrt(1.0 + R[1][1] - R[0][0] - R[2][2]) * 2; q[0] = (R[0][2] - R[2][0]) / s; q[1] = (R[0][1] + R[1][0]) / s; q[2] = 0.25 * s; q[3] = (R[1][2] + R[2][1]) / s; }
  else { double s = sqrt(1.0 + R[2][2] - R[0][0] - R[1][1]) * 2; q[0] = (R[1][0] - R[0][1]) / s; q[1] = (R[0][2] + R[2][0]) / s; q[2] = (R[1][2] + R[2][1]) / s; q[3] = 0.25 * s; }
  double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] /= nrm;
}

/* ---------------- configuration ---------------- */
int orc_set_box(orc_ctx *c, const double lo[3], const double hi[3], const int periodic[3]) {
  for (int k = 0; k < 3; k++) { c->lo[k] = lo[k]; c->hi[k] = hi[k]; c->periodic[k] = periodic[k]; if (!(hi[k] > lo[k])) return fail(c, "box: hi <= lo"); }
  c->forces_valid = 0; return 0;
}
int orc_set_quadrature(orc_ctx *c, int n_theta, int n_phi) {
  if (c->nshape) return fail(c, "set_quadrature must precede add_shape");
  if (n_theta < 2 || n_phi < 4) return fail(c, "quadrature too small");
  c->n_theta = n_theta; c->n_phi = n_phi; return 0;
}

/* folded recurrence tables (SURVEY App. B "folded form") */
static void build_folded(shape_t *s) {
  int L = s->lmax, T = s->T;
  s->off = (int *)malloc((L + 2) * sizeof(int));
  s->Ap = (double *)calloc(T, sizeof(double)); s->ah = (double *)calloc(T, sizeof(double)); s->bh = (double *)calloc(T, sizeof(double));
  double *alpha = (double *)malloc((L + 1) * sizeof(double));
  int o = 0; double cm = sqrt(1.0 / (4.0 * PI));
  for (int m = 0; m <= L; m++) {
    s->off[m] = o;
    if (m > 0) cm = cm * sqrt((2.0 * m + 1.0) / (2.0 * m));
    for (int l = m; l <= L; l++) {
      int idx = o + (l - m), k = l * (l + 1) / 2 + m;
      double Ap;
      if (l == m) { alpha[l] = 1.0; Ap = 0.0; }
      else if (l == m + 1) { alpha[l] = 1.0; Ap = sqrt(2.0 * m + 3.0); }
      else {
        double A = sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
        double B = sqrt((((double)l - 1.0) * (l - 1.0) - (double)m * m) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
        alpha[l] = (A * B) * alpha[l - 2];
        Ap = (A * alpha[l - 1]) / alpha[l];
      }
      s->Ap[idx] = Ap;
      s->ah[idx] = (s->a[k] * alpha[l]) * cm;
      s->bh[idx] = (m == 0) ? 0.0 : (s->b[k] * alpha[l]) * cm;
    }
    o += L + 1 - m;
  }
  s->off[L + 1] = o;
  free(alpha);
}

/* hot-path radius evaluation: folded recurrences, trig-free (A.2).
 * (sx,sy,sz) is the point in the shape frame, rho2 = |s|^2.  Returns r(dir s); *rho_out = |s|. */
static inline double sh_radius_folded(const shape_t *s, double sx, double sy, double sz, double rho2, double *rho_out) {
  const int L = s->lmax;
  const double *Ap = s->Ap, *ah = s->ah, *bh = s->bh;
  double rho = sqrt(rho2), inv = 1.0 / rho;
  double x = sz * inv, zx = sx * inv, zy = sy * inv;
  double u = 1.0, v = 0.0, r = 0.0;
  int base = 0;
  for (int m = 0; m <= L; m++) {
    if (m > 0) {
      double t1 = v * zy, un = fma(u, zx, -t1);
      double t2 = v * zx, vn = fma(u, zy, t2);
      u = un; v = vn;
    }
    double C = ah[base], S = bh[base];
    if (m < L) {
      double q1 = Ap[base + 1] * x, q2 = 1.0;
      C = fma(ah[base + 1], q1, C); S = fma(bh[base + 1], q1, S);
      for (int i = 2; i <= L - m; i++) {
        double tx = Ap[base + i] * x;
        double q = fma(tx, q1, -q2);
        C = fma(ah[base + i], q, C); S = fma(bh[base + i], q, S);
        q2 = q1; q1 = q;
      }
    }
    r = fma(u, C, r); r = fma(v, S, r);
    base += L + 1 - m;
  }
  *rho_out = rho;
  return r;
}

int orc_add_shape(orc_ctx *c, int lmax, const double *a_lm, const double *b_lm, double density, int *shape_id_out) {
  if (c->nshape >= MAX_SHAPES) return fail(c, "too many shapes");
  if (lmax < 0 || lmax > 128) return fail(c, "lmax out of range");
  if (!(density > 0)) return fail(c, "density must be > 0");
  shape_t *s = &c->shp[c->nshape]; memset(s, 0, sizeof *s);
  int L = lmax, T = (L + 1) * (L + 2) / 2;
  s->lmax = L; s->T = T; s->density = density;
  s->a = (double *)malloc(T * sizeof(double)); s->b = (double *)malloc(T * sizeof(double));
  memcpy(s->a, a_lm, T * sizeof(double));
  if (b_lm) memcpy(s->b, b_lm, T * sizeof(double)); else memset(s->b, 0, T * sizeof(double));
  for (int l = 0; l <= L; l++) s->b[l * (l + 1) / 2] = 0.0;
  build_folded(s);

  int nt = c->n_theta, np = c->n_phi, nq = nt * np;
  s->nq = nq; s->p = (double *)malloc(nq * 3 * sizeof(double)); s->nds = (double *)malloc(nq * 3 * sizeof(double));
  double *gx = (double *)malloc(nt * sizeof(double)), *gw = (double *)malloc(nt * sizeof(double));
  double *P = (double *)malloc(T * sizeof(double));
  orc_gauss_legendre(nt, gx, gw);
  double dphi = 2.0 * PI / np;
  double vol = 0, m1[3] = {0, 0, 0}, Io[3][3] = {{0}};
  double rmax = 0, rmin = 1e300;
  for (int a = 0; a < nt; a++) {
    double theta = acos(gx[a]);
    for (int b = 0; b < np; b++) {
      double phi = (b + 0.5) * dphi, r, rth, rph;
      shape_eval_setup(s, P, theta, phi, &r, &rth, &rph);
      if (!(r > 0)) { free(gx); free(gw); free(P); free_shape(s); return fail(c, "shape not star-shaped: r <= 0 at a node"); }
      double st = sin(theta), ct = cos(theta), cp = cos(phi), sp = sin(phi);
      double rh[3] = {st * cp, st * sp, ct}, th[3] = {ct * cp, ct * sp, -st}, ph[3] = {-sp, cp, 0.0};
      double w = gw[a] * dphi;
      int k = a * np + b;
      for (int d = 0; d < 3; d++) {
        s->p[3 * k + d] = r * rh[d];
        s->nds[3 * k + d] = (r * r * rh[d] - r * rth * th[d] - (r * rph / st) * ph[d]) * w;
      }
      double r3 = r * r * r;
      vol += w * r3 / 3.0;
      for (int d = 0; d < 3; d++) m1[d] += w * (r3 * r / 4.0) * rh[d];
      double r5 = r3 * r * r / 5.0;
      for (int d = 0; d < 3; d++) for (int e = 0; e < 3; e++) Io[d][e] += w * r5 * ((d == e) - rh[d] * rh[e]);
      if (r > rmax) rmax = r;
      if (r < rmin) rmin = r;
    }
  }
  /* dense sampling for the bounding radii (A.3 "refined") */
  int dt_ = 4 * nt, dp_ = 4 * np;
  for (int a = 0; a < dt_; a++)
    for (int b = 0; b < dp_; b++) {
      double r, rth, rph;
      shape_eval_setup(s, P, (a + 0.5) * PI / dt_, (b + 0.5) * 2.0 * PI / dp_, &r, &rth, &rph);
      if (!(r > 0)) { free(gx); free(gw); free(P); free_shape(s); return fail(c, "shape not star-shaped: r <= 0"); }
      if (r > rmax) rmax = r;
      if (r < rmin) rmin = r;
    }
  s->rmax = 1.005 * rmax; s->rmin = 0.995 * rmin;
  s->volume = vol; s->mass = density * vol;
  for (int d = 0; d < 3; d++) s->com[d] = m1[d] / vol;
  double Ic[3][3];
  double c2 = s->com[0] * s->com[0] + s->com[1] * s->com[1] + s->com[2] * s->com[2];
  for (int d = 0; d < 3; d++) for (int e = 0; e < 3; e++)
    Ic[d][e] = density * Io[d][e] - s->mass * (c2 * (d == e) - s->com[d] * s->com[e]);
  for (int d = 0; d < 3; d++) for (int e = d + 1; e < 3; e++) { double av = 0.5 * (Ic[d][e] + Ic[e][d]); Ic[d][e] = Ic[e][d] = av; }
  double ev[3], V[3][3];
  jacobi3(Ic, ev, V);
  /* right-handed principal frame: third axis = e0 x e1 */
  V[0][2] = V[1][0] * V[2][1] - V[2][0] * V[1][1];
  V[1][2] = V[2][0] * V[0][1] - V[0][0] * V[2][1];
  V[2][2] = V[0][0] * V[1][1] - V[1][0] * V[0][1];
  for (int d = 0; d < 3; d++) s->inertia[d] = ev[d];
  mat_to_quat(V, s->qp);
  quat_to_mat(s->qp, s->Rp);
  free(gx); free(gw); free(P);
  if (shape_id_out) *shape_id_out = c->nshape;
  c->nshape++;
  c->forces_valid = 0;
  return 0;
}

int orc_get_shape_props(const orc_ctx *c, int shape, double *volume, double com[3], double inertia[3],
                        double quat_principal[4], double *rmax, double *rmin) {
  if (shape < 0 || shape >= c->nshape) return -1;
  const shape_t *s = &c->shp[shape];
  if (volume) *volume = s->volume;
  if (com) memcpy(com, s->com, 3 * sizeof(double));
  if (inertia) memcpy(inertia, s->inertia, 3 * sizeof(double));
  if (quat_principal) memcpy(quat_principal, s->qp, 4 * sizeof(double));
  if (rmax) *rmax = s->rmax;
  if (rmin) *rmin = s->rmin;
  return 0;
}
int orc_get_nodes(const orc_ctx *c, int shape, double *p, double *nds) {
  if (shape < 0 || shape >= c->nshape) return -1;
  const shape_t *s = &c->shp[shape];
  memcpy(p, s->p, s->nq * 3 * sizeof(double)); memcpy(nds, s->nds, s->nq * 3 * sizeof(double));
  return 0;
}

int orc_set_atoms(orc_ctx *c, int64_t n, const int64_t *tag, const int *shape, const double *x,
                  const double *v, const double *quat, const double *angmom) {
  if (n < 0) return fail(c, "n < 0");
  for (int64_t i = 0; i < n; i++) if (shape[i] < 0 || shape[i] >= c->nshape) return fail(c, "atom shape id out of range");
  free(c->tag); free(c->shape); free(c->x); free(c->v); free(c->q); free(c->L); free(c->f); free(c->tq); free(c->Rs); free(c->c); free(c->c0);
  c->n = n; c->nghost = 0;
  size_t m = (size_t)(n > 0 ? n : 1);
  c->tag = (int64_t *)malloc(m * sizeof(int64_t)); c->shape = (int *)malloc(m * sizeof(int));
  c->x = (double *)calloc(m * 3, 8); c->v = (double *)calloc(m * 3, 8); c->q = (double *)calloc(m * 4, 8);
  c->L = (double *)calloc(m * 3, 8); c->f = (double *)calloc(m * 3, 8); c->tq = (double *)calloc(m * 3, 8);
  c->Rs = (double *)calloc(m * 9, 8); c->c = (double *)calloc(m * 3, 8); c->c0 = (double *)calloc(m * 3, 8);
  for (int64_t i = 0; i < n; i++) {
    c->tag[i] = tag ? tag[i] : i + 1; c->shape[i] = shape[i];
    for (int d = 0; d < 3; d++) { c->x[3 * i + d] = x[3 * i + d]; c->v[3 * i + d] = v ? v[3 * i + d] : 0.0; c->L[3 * i + d] = angmom ? angmom[3 * i + d] : 0.0; }
    if (quat) {
      double nn = sqrt(quat[4 * i] * quat[4 * i] + quat[4 * i + 1] * quat[4 * i + 1] + quat[4 * i + 2] * quat[4 * i + 2] + quat[4 * i + 3] * quat[4 * i + 3]);
      if (!(nn > 0)) return fail(c, "zero quaternion");
      for (int d = 0; d < 4; d++) c->q[4 * i + d] = quat[4 * i + d] / nn;
    } else { c->q[4 * i] = 1.0; }
  }
  c->forces_valid = 0; c->npair = 0;
  return 0;
}
int orc_pair_coeff(orc_ctx *c, int si, int sj, double k, double m) {
  if (si < 0 || sj < 0 || si >= MAX_SHAPES || sj >= MAX_SHAPES) return fail(c, "pair_coeff: shape out of range");
  if (!(k >= 0) || !(m >= 1.0)) return fail(c, "pair_coeff: need k >= 0, exponent >= 1");
  c->pk[si][sj] = c->pk[sj][si] = k; c->pm[si][sj] = c->pm[sj][si] = m; c->forces_valid = 0; return 0;
}
int orc_pair_dissipation(orc_ctx *c, int si, int sj, double gamma_n, double gamma_t, double mu) {
  if (si < 0 || sj < 0 || si >= MAX_SHAPES || sj >= MAX_SHAPES) return fail(c, "pair_dissipation: shape out of range");
  if (gamma_n < 0 || gamma_t < 0 || mu < 0) return fail(c, "pair_dissipation: negative coefficient");
  c->gn[si][sj] = c->gn[sj][si] = gamma_n; c->gt[si][sj] = c->gt[sj][si] = gamma_t; c->mu[si][sj] = c->mu[sj][si] = mu;
  c->forces_valid = 0; return 0;
}
int orc_add_wall(orc_ctx *c, const double point[3], const double normal[3], double k, double m) {
  if (c->nwall >= MAX_WALLS) return fail(c, "too many walls");
  double nn = sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
  if (!(nn > 0)) return fail(c, "wall normal is zero");
  if (!(k >= 0) || !(m >= 1.0)) return fail(c, "wall: need k >= 0, exponent >= 1");
  wall_t *w = &c->wall[c->nwall++];
  for (int d = 0; d < 3; d++) { w->c[d] = point[d]; w->n[d] = normal[d] / nn; }
  w->k = k; w->m = m; c->forces_valid = 0; return 0;
}
int orc_set_gravity(orc_ctx *c, const double g[3]) { for (int d = 0; d < 3; d++) c->g[d] = g[d]; return 0; }
int orc_set_neighbor(orc_ctx *c, double skin, int every, int check) { (void)every; (void)check; if (skin < 0) return fail(c, "skin < 0"); c->skin = skin; return 0; }
int orc_set_timestep(orc_ctx *c, double dt) { if (!(dt > 0)) return fail(c, "dt <= 0"); c->dt = dt; return 0; }
int orc_set_damping(orc_ctx *c, double gl, double gr) { if (gl < 0 || gr < 0) return fail(c, "damping < 0"); c->gamma_lin = gl; c->gamma_rot = gr; return 0; }
int orc_set_shear(orc_ctx *c, double rate) {
  if (rate != 0.0 && !(c->periodic[0] && c->periodic[1])) return fail(c, "shear needs a box periodic in x and y");
  c->shear_rate = rate; c->time = 0.0; c->forces_valid = 0;
  return 0;
}
int orc_set_threads(orc_ctx *c, int nthreads) { c->nthreads = nthreads > 0 ? nthreads : 1; return 0; }

/* ---------------- pose: Rs = R(q) Rp^T, c = x - Rs com ---------------- */
static void compute_pose(orc_ctx *c) {
#pragma omp parallel for num_threads(c->nthreads)
  for (int64_t i = 0; i < c->n; i++) {
    const shape_t *s = &c->shp[c->shape[i]];
    double Rq[3][3]; quat_to_mat(&c->q[4 * i], Rq);
    double *Rs = &c->Rs[9 * i];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) {
        double m = Rq[r][0] * s->Rp[k][0];
        m = fma(Rq[r][1], s->Rp[k][1], m);
        m = fma(Rq[r][2], s->Rp[k][2], m);
        Rs[3 * r + k] = m;
      }
    for (int r = 0; r < 3; r++) {
      double t = Rs[3 * r] * s->com[0];
      t = fma(Rs[3 * r + 1], s->com[1], t);
      t = fma(Rs[3 * r + 2], s->com[2], t);
      c->c[3 * i + r] = c->x[3 * i + r] - t;
    }
  }
}

/* ---------------- neighbor list (A.8): unordered pairs i<j, bounding spheres + skin ---------------- */
static void min_image(const orc_ctx *c, double d[3]) {
  if (c->shear_rate != 0.0) {
    /* Lees-Edwards: the image cell n_y boxes up is displaced by n_y * offset(t) along x (SURVEY §8d cfg 4).  Coordinates
       are never wrapped here, so the offset is simply rate * Ly * t. */
    const double Lx = c->hi[0] - c->lo[0], Ly = c->hi[1] - c->lo[1], off = c->shear_rate * Ly * c->time;
    const double ny = rint(d[1] / Ly);
    d[1] = d[1] - Ly * ny;
    d[0] = d[0] - ny * off;
    d[0] = d[0] - Lx * rint(d[0] / Lx);
    if (c->periodic[2]) { double Lz = c->hi[2] - c->lo[2]; d[2] = d[2] - Lz * rint(d[2] / Lz); }
    return;
  }
  for (int k = 0; k < 3; k++)
    if (c->periodic[k]) { double Lk = c->hi[k] - c->lo[k]; d[k] = d[k] - Lk * rint(d[k] / Lk); }
}
static void push_pair(orc_ctx *c, int i, int j) {
  if (c->npair == c->cappair) { c->cappair = c->cappair ? 2 * c->cappair : 1024; c->pr = (pairres_t *)realloc(c->pr, c->cappair * sizeof(pairres_t)); }
  memset(&c->pr[c->npair], 0, sizeof(pairres_t));
  c->pr[c->npair].i = i; c->pr[c->npair].j = j; c->npair++;
}
static int pair_cmp(const void *a, const void *b) {
  const pairres_t *p = (const pairres_t *)a, *q = (const pairres_t *)b;
  if (p->i != q->i) return p->i < q->i ? -1 : 1;
  return p->j < q->j ? -1 : (p->j > q->j);
}
static int build_neighbors(orc_ctx *c) {
  c->npair = 0;
  int64_t n = c->n;
  double rmaxg = 0;
  for (int s = 0; s < c->nshape; s++) if (c->shp[s].rmax > rmaxg) rmaxg = c->shp[s].rmax;
  double cut = 2.0 * rmaxg + c->skin;
  for (int k = 0; k < 3; k++) if (c->periodic[k] && (c->hi[k] - c->lo[k]) < 2.0 * cut) return fail(c, "periodic box shorter than 2x cutoff");
  if (n <= 2000) {
    for (int64_t i = 0; i < n; i++)
      for (int64_t j = i + 1; j < n; j++) {
        double d[3] = {c->c[3 * i] - c->c[3 * j], c->c[3 * i + 1] - c->c[3 * j + 1], c->c[3 * i + 2] - c->c[3 * j + 2]};
        min_image(c, d);
        double rc = c->shp[c->shape[i]].rmax + c->shp[c->shape[j]].rmax + c->skin;
        if (i >= c->n - c->nghost) continue;   /* ghost-ghost pairs belong to other ranks */
        if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] < rc * rc) push_pair(c, (int)i, (int)j);
      }
    return 0;
  }
  /* cell list */
  double lo[3], hi[3]; int nc[3];
  for (int k = 0; k < 3; k++) {
    if (c->periodic[k]) { lo[k] = c->lo[k]; hi[k] = c->hi[k]; }
    else { lo[k] = 1e300; hi[k] = -1e300; for (int64_t i = 0; i < n; i++) { double v = c->c[3 * i + k]; if (v < lo[k]) lo[k] = v; if (v > hi[k]) hi[k] = v; } hi[k] += 1e-9 * (1 + fabs(hi[k])); }
    nc[k] = (int)floor((hi[k] - lo[k]) / cut); if (nc[k] < 1) nc[k] = 1; if (nc[k] > 512) nc[k] = 512;
    if (c->periodic[k] && nc[k] < 3) nc[k] = 1;
  }
  int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
  int *head = (int *)malloc(ncell * sizeof(int)), *next = (int *)malloc(n * sizeof(int)), *cell = (int *)malloc(n * 3 * sizeof(int));
  for (int64_t k = 0; k < ncell; k++) head[k] = -1;
  for (int64_t i = n - 1; i >= 0; i--) {
    int ci[3];
    for (int k = 0; k < 3; k++) {
      double Lk = hi[k] - lo[k], u = c->c[3 * i + k] - lo[k];
      if (c->periodic[k]) u -= Lk * floor(u / Lk);
      int b = (int)(u / Lk * nc[k]); if (b < 0) b = 0; if (b >= nc[k]) b = nc[k] - 1; ci[k] = b; cell[3 * i + k] = b;
    }
    int64_t id = ((int64_t)ci[2] * nc[1] + ci[1]) * nc[0] + ci[0];
    next[i] = head[id]; head[id] = (int)i;
  }
  for (int64_t i = 0; i < n; i++) {
    int lo3[3], hi3[3];
    for (int k = 0; k < 3; k++) {
      if (nc[k] == 1) { lo3[k] = 0; hi3[k] = 0; }
      else if (c->periodic[k]) { lo3[k] = cell[3 * i + k] - 1; hi3[k] = cell[3 * i + k] + 1; }
      else { lo3[k] = cell[3 * i + k] > 0 ? cell[3 * i + k] - 1 : 0; hi3[k] = cell[3 * i + k] < nc[k] - 1 ? cell[3 * i + k] + 1 : nc[k] - 1; }
    }
    for (int cz = lo3[2]; cz <= hi3[2]; cz++) for (int cy = lo3[1]; cy <= hi3[1]; cy++) {
      /* under shear the row across the y boundary is displaced along x by an arbitrary amount: scan the whole row */
      const int sheared_row = c->shear_rate != 0.0 && (cy < 0 || cy >= nc[1]) && nc[0] > 1;
      const int x0 = sheared_row ? 0 : lo3[0], x1 = sheared_row ? nc[0] - 1 : hi3[0];
      for (int cx = x0; cx <= x1; cx++) {
      int wx = (cx % nc[0] + nc[0]) % nc[0], wy = (cy % nc[1] + nc[1]) % nc[1], wz = (cz % nc[2] + nc[2]) % nc[2];
      int64_t id = ((int64_t)wz * nc[1] + wy) * nc[0] + wx;
      for (int j = head[id]; j >= 0; j = next[j]) {
        if (j <= i) continue;
        double d[3] = {c->c[3 * i] - c->c[3 * j], c->c[3 * i + 1] - c->c[3 * j + 1], c->c[3 * i + 2] - c->c[3 * j + 2]};
        min_image(c, d);
        double rc = c->shp[c->shape[i]].rmax + c->shp[c->shape[j]].rmax + c->skin;
        if (i >= c->n - c->nghost) continue;
        if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] < rc * rc) push_pair(c, (int)i, j);
      }
      }
    }
  }
  free(head); free(next); free(cell);
  qsort(c->pr, c->npair, sizeof(pairres_t), pair_cmp);
  return 0;
}

/* ---------------- one direction of a pair (A.4): nodes of a against the surface of b ---------------- */
typedef struct { double S[3], A, T[3], G[3]; int64_t ntrans, neval, ninside; } dirsum_t;

static void eval_direction(const shape_t *sa, const shape_t *sb, const double *Ra, const double *Rb,
                           const double d[3] /* c_a - c_b, min image */, double halfsign, dirsum_t *out) {
  /* M = Rb^T Ra, t = Rb^T d ; x0 (midpoint) in a's frame = Ra^T (-0.5 d) */
  double M[3][3], t[3], x0a[3];
  for (int r = 0; r < 3; r++) {
    for (int k = 0; k < 3; k++) {
      double m = Rb[0 + r] * Ra[0 + k];
      m = fma(Rb[3 + r], Ra[3 + k], m);
      m = fma(Rb[6 + r], Ra[6 + k], m);
      M[r][k] = m;
    }
    double tt = Rb[0 + r] * d[0]; tt = fma(Rb[3 + r], d[1], tt); tt = fma(Rb[6 + r], d[2], tt); t[r] = tt;
    double hx = Ra[0 + r] * d[0]; hx = fma(Ra[3 + r], d[1], hx); hx = fma(Ra[6 + r], d[2], hx); x0a[r] = halfsign * hx;
  }
  (void)sa;
  double rmax2 = sb->rmax * sb->rmax, rmin2 = sb->rmin * sb->rmin;
  double S[3] = {0, 0, 0}, A = 0, T[3] = {0, 0, 0}, G[3] = {0, 0, 0};
  int64_t neval = 0, nin = 0;
  const int nq = sa->nq;
  for (int k = 0; k < nq; k++) {
    const double *p = &sa->p[3 * k];
    double s0 = fma(M[0][0], p[0], t[0]); s0 = fma(M[0][1], p[1], s0); s0 = fma(M[0][2], p[2], s0);
    double s1 = fma(M[1][0], p[0], t[1]); s1 = fma(M[1][1], p[1], s1); s1 = fma(M[1][2], p[2], s1);
    double s2 = fma(M[2][0], p[0], t[2]); s2 = fma(M[2][1], p[1], s2); s2 = fma(M[2][2], p[2], s2);
    double rho2 = fma(s2, s2, fma(s1, s1, s0 * s0));
    if (rho2 >= rmax2) continue;
    int inside;
    if (rho2 <= rmin2) inside = 1;
    else { double rho, r = sh_radius_folded(sb, s0, s1, s2, rho2, &rho); inside = rho < r; neval++; }
    if (!inside) continue;
    nin++;
    const double *n = &sa->nds[3 * k];
    double dp0 = p[0] - x0a[0], dp1 = p[1] - x0a[1], dp2 = p[2] - x0a[2];
    double dn = fma(dp2, n[2], fma(dp1, n[1], dp0 * n[0]));
    S[0] += n[0]; S[1] += n[1]; S[2] += n[2];
    A += dn;
    T[0] += fma(p[1], n[2], -(p[2] * n[1]));
    T[1] += fma(p[2], n[0], -(p[0] * n[2]));
    T[2] += fma(p[0], n[1], -(p[1] * n[0]));
    G[0] = fma(dp0, dn, G[0]); G[1] = fma(dp1, dn, G[1]); G[2] = fma(dp2, dn, G[2]);
  }
  /* rotate body-frame sums to the space frame */
  for (int r = 0; r < 3; r++) {
    out->S[r] = Ra[3 * r] * S[0] + Ra[3 * r + 1] * S[1] + Ra[3 * r + 2] * S[2];
    out->T[r] = Ra[3 * r] * T[0] + Ra[3 * r + 1] * T[1] + Ra[3 * r + 2] * T[2];
    out->G[r] = 0.25 * (Ra[3 * r] * G[0] + Ra[3 * r + 1] * G[1] + Ra[3 * r + 2] * G[2]);
  }
  out->A = A / 3.0;
  out->ntrans = nq; out->neval = neval; out->ninside = nin;
}

static double contact_pressure(double k, double m, double V, double *E) {
  if (m == 1.0) { *E = k * V; return k; }
  double pw = pow(V, m - 1.0);
  *E = k * pw * V;
  return m * k * pw;
}

static void omega_from_L(const double q[4], const double L[3], const double I[3], double w[3]);

/* A.5b  Dissipative part of the contact law (builder's choice; the reference's is unknown: PARITY UNPINNED).
 * History-free viscous normal damping + regularised Coulomb friction, applied at the overlap centroid x_c:
 *   n = F_el / |F_el|  (direction of the elastic force on i),  r_i = x_c - x_i, r_j = x_c - x_j,
 *   v_rel = (v_i + w_i x r_i) - (v_j + w_j x r_j),  vn = v_rel . n,
 *   normal magnitude  fnt = max(0, |F_el| - gamma_n vn)   (never attractive),
 *   tangential        F_t = -min(gamma_t |v_t|, mu fnt) v_t / |v_t|,  v_t = v_rel - vn n,
 *   F_i = fnt n + F_t = -F_j;  torques  tau_i += r_i x (F_i - F_el),  tau_j -= r_j x (F_i - F_el).
 * Velocities and angular momenta are the arrays as they stand when the forces are computed (half-step values inside
 * the velocity-Verlet step).  xj = centre of mass of j in i's periodic image. */
static void contact_dissipation(double gn, double gt, double mu, const double xi[3], const double xj[3], const double xc[3],
                                const double vi[3], const double vj[3], const double wi[3], const double wj[3],
                                double F[3], double ti[3], double tj[3]) {
  const double fn2 = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
  if (!(fn2 > 0.0)) return;
  const double fn = sqrt(fn2);
  const double nh[3] = {F[0] / fn, F[1] / fn, F[2] / fn};
  const double ri[3] = {xc[0] - xi[0], xc[1] - xi[1], xc[2] - xi[2]}, rj[3] = {xc[0] - xj[0], xc[1] - xj[1], xc[2] - xj[2]};
  const double vr[3] = {(vi[0] + (wi[1] * ri[2] - wi[2] * ri[1])) - (vj[0] + (wj[1] * rj[2] - wj[2] * rj[1])),
                        (vi[1] + (wi[2] * ri[0] - wi[0] * ri[2])) - (vj[1] + (wj[2] * rj[0] - wj[0] * rj[2])),
                        (vi[2] + (wi[0] * ri[1] - wi[1] * ri[0])) - (vj[2] + (wj[0] * rj[1] - wj[1] * rj[0]))};
  const double vn = vr[0] * nh[0] + vr[1] * nh[1] + vr[2] * nh[2];
  double fnt = fn - gn * vn;
  if (fnt < 0.0) fnt = 0.0;
  double Fd[3] = {(fnt - fn) * nh[0], (fnt - fn) * nh[1], (fnt - fn) * nh[2]};
  const double vt[3] = {vr[0] - vn * nh[0], vr[1] - vn * nh[1], vr[2] - vn * nh[2]};
  const double vt2 = vt[0] * vt[0] + vt[1] * vt[1] + vt[2] * vt[2];
  if (gt > 0.0 && mu > 0.0 && vt2 > 0.0) {
    const double vtm = sqrt(vt2);
    double ft = gt * vtm;
    const double cap = mu * fnt;
    if (ft > cap) ft = cap;
    const double sc = ft / vtm;
    Fd[0] -= sc * vt[0]; Fd[1] -= sc * vt[1]; Fd[2] -= sc * vt[2];
  }
  for (int r = 0; r < 3; r++) F[r] += Fd[r];
  ti[0] += ri[1] * Fd[2] - ri[2] * Fd[1]; ti[1] += ri[2] * Fd[0] - ri[0] * Fd[2]; ti[2] += ri[0] * Fd[1] - ri[1] * Fd[0];
  tj[0] -= rj[1] * Fd[2] - rj[2] * Fd[1]; tj[1] -= rj[2] * Fd[0] - rj[0] * Fd[2]; tj[2] -= rj[0] * Fd[1] - rj[1] * Fd[0];
}

/* ---------------- forces: pairs (A.4, A.5), walls (A.6), gather ---------------- */
int orc_compute_forces(orc_ctx *c) {
  if (c->n > 0 && c->nshape == 0) return fail(c, "no shapes");
  compute_pose(c);
  if (build_neighbors(c)) return -1;
  int64_t ctr = 0, cev = 0, cin = 0;
#pragma omp parallel for schedule(dynamic, 4) num_threads(c->nthreads) reduction(+ : ctr, cev, cin)
  for (int64_t pidx = 0; pidx < c->npair; pidx++) {
    pairres_t *pr = &c->pr[pidx];
    int i = pr->i, j = pr->j;
    const shape_t *si = &c->shp[c->shape[i]], *sj = &c->shp[c->shape[j]];
    const double *Ri = &c->Rs[9 * i], *Rj = &c->Rs[9 * j];
    double d[3] = {c->c[3 * i] - c->c[3 * j], c->c[3 * i + 1] - c->c[3 * j + 1], c->c[3 * i + 2] - c->c[3 * j + 2]};
    min_image(c, d);
    double dm[3] = {-d[0], -d[1], -d[2]};
    dirsum_t ij, ji;
    eval_direction(si, sj, Ri, Rj, d, -0.5, &ij);  /* nodes of i in j; x0 = c_i - d/2 */
    eval_direction(sj, si, Rj, Ri, dm, -0.5, &ji); /* nodes of j in i; x0 = c_j + d/2 */
    ctr += ij.ntrans + ji.ntrans; cev += ij.neval + ji.neval; cin += ij.ninside + ji.ninside;
    pr->ninside = (int)(ij.ninside + ji.ninside);
    double V = ij.A + ji.A;
    pr->V = 0; pr->E = 0;
    for (int r = 0; r < 3; r++) { pr->F[r] = 0; pr->ti[r] = 0; pr->tj[r] = 0; pr->xc[r] = 0; }
    if (pr->ninside == 0 || !(V > 0)) continue;
    double E, p = contact_pressure(c->pk[c->shape[i]][c->shape[j]], c->pm[c->shape[i]][c->shape[j]], V, &E);
    pr->V = V; pr->E = E;
    double Fi[3];
    for (int r = 0; r < 3; r++) { Fi[r] = -p * (0.5 * (ij.S[r] - ji.S[r])); pr->F[r] = Fi[r]; }
    /* torques about each particle's centre of mass x = c + Rs com: lever (c - x) x (own-surface force) */
    double li[3], lj[3];
    for (int r = 0; r < 3; r++) { li[r] = c->c[3 * i + r] - c->x[3 * i + r]; lj[r] = c->c[3 * j + r] - c->x[3 * j + r]; }
    double Ti[3] = {ij.T[0] + (li[1] * ij.S[2] - li[2] * ij.S[1]), ij.T[1] + (li[2] * ij.S[0] - li[0] * ij.S[2]), ij.T[2] + (li[0] * ij.S[1] - li[1] * ij.S[0])};
    double Tj[3] = {ji.T[0] + (lj[1] * ji.S[2] - lj[2] * ji.S[1]), ji.T[1] + (lj[2] * ji.S[0] - lj[0] * ji.S[2]), ji.T[2] + (lj[0] * ji.S[1] - lj[1] * ji.S[0])};
    for (int r = 0; r < 3; r++) {
      pr->ti[r] = -p * Ti[r]; pr->tj[r] = -p * Tj[r];
      pr->xc[r] = (c->c[3 * i + r] - 0.5 * d[r]) + (ij.G[r] + ji.G[r]) / V;
    }
    const double gn = c->gn[c->shape[i]][c->shape[j]], gt = c->gt[c->shape[i]][c->shape[j]], mu = c->mu[c->shape[i]][c->shape[j]];
    if (gn > 0.0 || (gt > 0.0 && mu > 0.0)) {
      double xj[3], wi[3], wj[3];
      for (int r = 0; r < 3; r++) xj[r] = (c->c[3 * i + r] - d[r]) - lj[r];       /* j's centre of mass in i's image */
      omega_from_L(&c->q[4 * i], &c->L[3 * i], si->inertia, wi);
      omega_from_L(&c->q[4 * j], &c->L[3 * j], sj->inertia, wj);
      double vj[3] = {c->v[3 * j], c->v[3 * j + 1], c->v[3 * j + 2]};
      if (c->shear_rate != 0.0) {   /* the image of j that i touches moves with the sheared cell it sits in */
        const double Ly = c->hi[1] - c->lo[1];
        const double ny = rint((c->c[3 * i + 1] - c->c[3 * j + 1]) / Ly);
        vj[0] += ny * c->shear_rate * Ly;
      }
      contact_dissipation(gn, gt, mu, &c->x[3 * i], xj, pr->xc, &c->v[3 * i], vj, wi, wj, pr->F, pr->ti, pr->tj);
    }
  }
  c->cnt_pairs += c->npair; c->cnt_trans += ctr; c->cnt_eval += cev; c->cnt_inside += cin;
  /* gather in pair order */
  memset(c->f, 0, (size_t)c->n * 3 * 8); memset(c->tq, 0, (size_t)c->n * 3 * 8);
  double ec = 0;
  for (int64_t pidx = 0; pidx < c->npair; pidx++) {
    const pairres_t *pr = &c->pr[pidx];
    for (int r = 0; r < 3; r++) {
      c->f[3 * pr->i + r] += pr->F[r]; c->f[3 * pr->j + r] -= pr->F[r];
      c->tq[3 * pr->i + r] += pr->ti[r]; c->tq[3 * pr->j + r] += pr->tj[r];
    }
    ec += pr->E;
  }
  /* walls (A.6) */
  for (int w = 0; w < c->nwall; w++) {
    const wall_t *wl = &c->wall[w];
    for (int64_t i = 0; i < c->n; i++) {
      const shape_t *s = &c->shp[c->shape[i]];
      const double *R = &c->Rs[9 * i];
      double dc[3] = {c->c[3 * i] - wl->c[0], c->c[3 * i + 1] - wl->c[1], c->c[3 * i + 2] - wl->c[2]};
      double h = fma(dc[2], wl->n[2], fma(dc[1], wl->n[1], dc[0] * wl->n[0]));
      if (h >= s->rmax) continue;
      double nb[3];
      for (int r = 0; r < 3; r++) { double t = R[r] * wl->n[0]; t = fma(R[3 + r], wl->n[1], t); t = fma(R[6 + r], wl->n[2], t); nb[r] = t; }
      double x0[3] = {-h * nb[0], -h * nb[1], -h * nb[2]};
      double S[3] = {0, 0, 0}, A = 0, T[3] = {0, 0, 0}; int nin = 0;
      for (int k = 0; k < s->nq; k++) {
        const double *p = &s->p[3 * k], *n = &s->nds[3 * k];
        double g = fma(nb[2], p[2], fma(nb[1], p[1], fma(nb[0], p[0], h)));
        if (!(g < 0)) continue;
        nin++;
        double dp0 = p[0] - x0[0], dp1 = p[1] - x0[1], dp2 = p[2] - x0[2];
        A += fma(dp2, n[2], fma(dp1, n[1], dp0 * n[0]));
        S[0] += n[0]; S[1] += n[1]; S[2] += n[2];
        T[0] += fma(p[1], n[2], -(p[2] * n[1]));
        T[1] += fma(p[2], n[0], -(p[0] * n[2]));
        T[2] += fma(p[0], n[1], -(p[1] * n[0]));
      }
      double V = A / 3.0;
      if (nin == 0 || !(V > 0)) continue;
      double E, p = contact_pressure(wl->k, wl->m, V, &E);
      ec += E;
      double Ss[3], Ts[3], l[3];
      for (int r = 0; r < 3; r++) {
        Ss[r] = R[3 * r] * S[0] + R[3 * r + 1] * S[1] + R[3 * r + 2] * S[2];
        Ts[r] = R[3 * r] * T[0] + R[3 * r + 1] * T[1] + R[3 * r + 2] * T[2];
        l[r] = c->c[3 * i + r] - c->x[3 * i + r];
      }
      double Tt[3] = {Ts[0] + (l[1] * Ss[2] - l[2] * Ss[1]), Ts[1] + (l[2] * Ss[0] - l[0] * Ss[2]), Ts[2] + (l[0] * Ss[1] - l[1] * Ss[0])};
      for (int r = 0; r < 3; r++) { c->f[3 * i + r] += -p * Ss[r]; c->tq[3 * i + r] += -p * Tt[r]; }
    }
  }
  c->e_contact = ec;
  c->forces_valid = 1;
  return 0;
}

/* ---------------- integrator (A.7) ---------------- */
static void omega_from_L(const double q[4], const double L[3], const double I[3], double w[3]) {
  double R[3][3]; quat_to_mat(q, R);
  double wb[3];
  for (int k = 0; k < 3; k++) { double lb = R[0][k] * L[0] + R[1][k] * L[1] + R[2][k] * L[2]; wb[k] = (I[k] > 0) ? lb / I[k] : 0.0; }
  for (int r = 0; r < 3; r++) w[r] = R[r][0] * wb[0] + R[r][1] * wb[1] + R[r][2] * wb[2];
}
static void vecquat(const double w[3], const double q[4], double o[4]) {
  o[0] = -w[0] * q[1] - w[1] * q[2] - w[2] * q[3];
  o[1] = q[0] * w[0] + w[1] * q[3] - w[2] * q[2];
  o[2] = q[0] * w[1] + w[2] * q[1] - w[0] * q[3];
  o[3] = q[0] * w[2] + w[0] * q[2] - w[1] * q[1];
}
static void qnorm(double q[4]) {
  double n = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] *= n;
}
static void richardson(double q[4], const double L[3], const double I[3], double dtq) {
  double w[3], wq[4], qf[4], qh[4];
  omega_from_L(q, L, I, w); vecquat(w, q, wq);
  for (int k = 0; k < 4; k++) { qf[k] = q[k] + dtq * wq[k]; qh[k] = q[k] + 0.5 * dtq * wq[k]; }
  qnorm(qf); qnorm(qh);
  omega_from_L(qh, L, I, w); vecquat(w, qh, wq);
  for (int k = 0; k < 4; k++) qh[k] += 0.5 * dtq * wq[k];
  qnorm(qh);
  for (int k = 0; k < 4; k++) q[k] = 2.0 * qh[k] - qf[k];
  qnorm(q);
}

int orc_run(orc_ctx *c, int64_t nsteps) {
  if (!c->forces_valid) if (orc_compute_forces(c)) return -1;
  double dt = c->dt, dth = 0.5 * dt;
  const double damp_v = 1.0 - 0.5 * dt * c->gamma_lin, damp_L = 1.0 - 0.5 * dt * c->gamma_rot;
  for (int64_t step = 0; step < nsteps; step++) {
#pragma omp parallel for num_threads(c->nthreads)
    for (int64_t i = 0; i < c->n; i++) {
      const shape_t *s = &c->shp[c->shape[i]];
      double im = 1.0 / s->mass;
      for (int d = 0; d < 3; d++) {
        c->v[3 * i + d] = (c->v[3 * i + d] + dth * (c->f[3 * i + d] * im + c->g[d])) * damp_v;
        c->x[3 * i + d] += dt * c->v[3 * i + d];
        c->L[3 * i + d] = (c->L[3 * i + d] + dth * c->tq[3 * i + d]) * damp_L;
      }
      richardson(&c->q[4 * i], &c->L[3 * i], s->inertia, dth);
    }
    c->time += dt;
    if (orc_compute_forces(c)) return -1;
#pragma omp parallel for num_threads(c->nthreads)
    for (int64_t i = 0; i < c->n; i++) {
      const shape_t *s = &c->shp[c->shape[i]];
      double im = 1.0 / s->mass;
      for (int d = 0; d < 3; d++) {
        c->v[3 * i + d] = (c->v[3 * i + d] + dth * (c->f[3 * i + d] * im + c->g[d])) * damp_v;
        c->L[3 * i + d] = (c->L[3 * i + d] + dth * c->tq[3 * i + d]) * damp_L;
      }
    }
  }
  return 0;
}

/* ---- multi-rank test support: same split step / ghost pack-unpack surface as libshgpu (host pointers) ---- */
int orc_set_ghost_count(orc_ctx *c, int64_t nghost) {
  if (nghost < 0 || nghost > c->n) return fail(c, "ghost count out of range");
  c->nghost = nghost; c->forces_valid = 0;
  compute_pose(c);
  memcpy(c->c0, c->c, (size_t)c->n * 24);
  return 0;
}
int orc_step_begin(orc_ctx *c, int *rebuild_wanted) {
  if (!c->forces_valid) { if (orc_compute_forces(c)) return -1; memcpy(c->c0, c->c, (size_t)c->n * 24); }
  const int64_t nl = c->n - c->nghost;
  const double dt = c->dt, dth = 0.5 * dt;
  const double damp_v = 1.0 - 0.5 * dt * c->gamma_lin, damp_L = 1.0 - 0.5 * dt * c->gamma_rot;
  for (int64_t i = 0; i < nl; i++) {
    const shape_t *s = &c->shp[c->shape[i]];
    double im = 1.0 / s->mass;
    for (int d = 0; d < 3; d++) {
      c->v[3 * i + d] = (c->v[3 * i + d] + dth * (c->f[3 * i + d] * im + c->g[d])) * damp_v;
      c->x[3 * i + d] += dt * c->v[3 * i + d];
      c->L[3 * i + d] = (c->L[3 * i + d] + dth * c->tq[3 * i + d]) * damp_L;
    }
    richardson(&c->q[4 * i], &c->L[3 * i], s->inertia, dth);
  }
  compute_pose(c);
  int flag = 0;
  const double trig2 = 0.25 * c->skin * c->skin;
  for (int64_t i = 0; i < nl && !flag; i++) {
    double d2 = 0;
    for (int d = 0; d < 3; d++) { double dd = c->c[3 * i + d] - c->c0[3 * i + d]; d2 += dd * dd; }
    if (d2 > trig2) flag = 1;
  }
  if (rebuild_wanted) *rebuild_wanted = flag;
  return 0;
}
int orc_step_end(orc_ctx *c, int rebuild) {
  if (orc_compute_forces(c)) return -1;
  if (rebuild) memcpy(c->c0, c->c, (size_t)c->n * 24);
  const int64_t nl = c->n - c->nghost;
  const double dt = c->dt, dth = 0.5 * dt;
  const double damp_v = 1.0 - 0.5 * dt * c->gamma_lin, damp_L = 1.0 - 0.5 * dt * c->gamma_rot;
  for (int64_t i = 0; i < nl; i++) {
    const shape_t *s = &c->shp[c->shape[i]];
    double im = 1.0 / s->mass;
    for (int d = 0; d < 3; d++) {
      c->v[3 * i + d] = (c->v[3 * i + d] + dth * (c->f[3 * i + d] * im + c->g[d])) * damp_v;
      c->L[3 * i + d] = (c->L[3 * i + d] + dth * c->tq[3 * i + d]) * damp_L;
    }
  }
  return 0;
}
int orc_pack_atoms(const orc_ctx *c, int64_t m, const int *idx, const double *shift, double *out) {
  for (int64_t k = 0; k < m; k++) {
    const int i = idx[k];
    for (int d = 0; d < 3; d++) out[7 * k + d] = c->x[3 * i + d] + (shift ? shift[3 * k + d] : 0.0);
    for (int d = 0; d < 4; d++) out[7 * k + 3 + d] = c->q[4 * i + d];
  }
  return 0;
}
int orc_unpack_ghosts(orc_ctx *c, int64_t first, int64_t m, const double *in) {
  if (first < 0 || first + m > c->n) return fail(c, "unpack range out of bounds");
  for (int64_t k = 0; k < m; k++) {
    for (int d = 0; d < 3; d++) c->x[3 * (first + k) + d] = in[7 * k + d];
    for (int d = 0; d < 4; d++) c->q[4 * (first + k) + d] = in[7 * k + 3 + d];
  }
  return 0;
}

int orc_get_atoms(const orc_ctx *c, int64_t n, double *x, double *v, double *quat, double *angmom, double *f, double *torque) {
  if (n != c->n) return -1;
  if (x) memcpy(x, c->x, (size_t)n * 24); if (v) memcpy(v, c->v, (size_t)n * 24);
  if (quat) memcpy(quat, c->q, (size_t)n * 32); if (angmom) memcpy(angmom, c->L, (size_t)n * 24);
  if (f) memcpy(f, c->f, (size_t)n * 24); if (torque) memcpy(torque, c->tq, (size_t)n * 24);
  return 0;
}
int orc_get_pairs(const orc_ctx *c, int64_t cap, int64_t *npairs, int64_t *tag_i, int64_t *tag_j, double *V, double *F,
                  double *tau_i, double *tau_j, double *centroid) {
  if (npairs) *npairs = c->npair;
  int64_t m = c->npair < cap ? c->npair : cap;
  for (int64_t k = 0; k < m; k++) {
    const pairres_t *pr = &c->pr[k];
    if (tag_i) tag_i[k] = c->tag[pr->i]; if (tag_j) tag_j[k] = c->tag[pr->j];
    if (V) V[k] = pr->V;
    for (int r = 0; r < 3; r++) {
      if (F) F[3 * k + r] = pr->F[r]; if (tau_i) tau_i[3 * k + r] = pr->ti[r];
      if (tau_j) tau_j[3 * k + r] = pr->tj[r]; if (centroid) centroid[3 * k + r] = pr->xc[r];
    }
  }
  return 0;
}
int orc_get_counters(const orc_ctx *c, int64_t *pair_evals, int64_t *nodes_transformed, int64_t *nodes_evaluated, int64_t *nodes_inside) {
  if (pair_evals) *pair_evals = c->cnt_pairs; if (nodes_transformed) *nodes_transformed = c->cnt_trans;
  if (nodes_evaluated) *nodes_evaluated = c->cnt_eval; if (nodes_inside) *nodes_inside = c->cnt_inside;
  return 0;
}
int orc_get_energy(const orc_ctx *c, double *ke_trans, double *ke_rot, double *e_contact) {
  double kt = 0, kr = 0;
  for (int64_t i = 0; i < c->n; i++) {
    const shape_t *s = &c->shp[c->shape[i]];
    const double *v = &c->v[3 * i];
    kt += 0.5 * s->mass * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double w[3]; omega_from_L(&c->q[4 * i], &c->L[3 * i], s->inertia, w);
    kr += 0.5 * (w[0] * c->L[3 * i] + w[1] * c->L[3 * i + 1] + w[2] * c->L[3 * i + 2]);
  }
  if (ke_trans) *ke_trans = kt; if (ke_rot) *ke_rot = kr; if (e_contact) *e_contact = c->e_contact;
  return 0;
}

/* Pressure-tensor sums (compute pressure / stress/atom in LAMMPS terms): kinetic[3a+b] = sum_i m v_a v_b over owned
 * atoms, virial[3a+b] = sum_pairs (x_i - x_j)_a F_b with the minimum-image centre-of-mass separation and F = force on i;
 * a pair with a ghost counts half (the owner rank of the ghost adds the other half).  Walls are not included.
 * pressure tensor = (kinetic + virial) / volume. */
int orc_get_stress(const orc_ctx *c, double virial[9], double kinetic[9]) {
  double W[9] = {0}, K[9] = {0};
  const int64_t nown = c->n - c->nghost;
  for (int64_t i = 0; i < nown; i++) {
    const double m = c->shp[c->shape[i]].mass, *v = &c->v[3 * i];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) K[3 * a + b] += m * v[a] * v[b];
  }
  if (c->forces_valid)
    for (int64_t p = 0; p < c->npair; p++) {
      const pairres_t *pr = &c->pr[p];
      const int i = pr->i, j = pr->j;
      double d[3] = {c->c[3 * i] - c->c[3 * j], c->c[3 * i + 1] - c->c[3 * j + 1], c->c[3 * i + 2] - c->c[3 * j + 2]};
      min_image(c, d);
      const double w = j >= nown ? 0.5 : 1.0;
      for (int a = 0; a < 3; a++) {
        const double dx = d[a] - (c->c[3 * i + a] - c->x[3 * i + a]) + (c->c[3 * j + a] - c->x[3 * j + a]);   /* x_i - x_j */
        for (int b = 0; b < 3; b++) W[3 * a + b] += w * dx * pr->F[b];
      }
    }
  for (int k = 0; k < 9; k++) { if (virial) virial[k] = W[k]; if (kinetic) kinetic[k] = K[k]; }
  return 0;
}

/* ---------------- KAT hooks ---------------- */
int orc_sh_radius(const orc_ctx *c, int shape, int64_t n, const double *dirs, double *r) {
  if (shape < 0 || shape >= c->nshape) return -1;
  const shape_t *s = &c->shp[shape];
  for (int64_t k = 0; k < n; k++) {
    const double *d = &dirs[3 * k];
    double rho2 = fma(d[2], d[2], fma(d[1], d[1], d[0] * d[0])), rho;
    r[k] = sh_radius_folded(s, d[0], d[1], d[2], rho2, &rho);
  }
  return 0;
}

int orc_project_ellipsoid(int lmax, double a, double b, double cc, int n_theta, int n_phi, double *a_lm, double *b_lm) {
  int T = (lmax + 1) * (lmax + 2) / 2;
  double *gx = (double *)malloc(n_theta * 8), *gw = (double *)malloc(n_theta * 8), *P = (double *)malloc(T * 8);
  orc_gauss_legendre(n_theta, gx, gw);
  memset(a_lm, 0, T * 8); memset(b_lm, 0, T * 8);
  double dphi = 2.0 * PI / n_phi;
  for (int i = 0; i < n_theta; i++) {
    double x = gx[i], st = sqrt((1.0 - x) * (1.0 + x));
    orc_legendre_norm(lmax, x, P);
    for (int j = 0; j < n_phi; j++) {
      double phi = (j + 0.5) * dphi, cp = cos(phi), sp = sin(phi);
      double ux = st * cp / a, uy = st * sp / b, uz = x / cc;
      double r = 1.0 / sqrt(ux * ux + uy * uy + uz * uz);
      double w = gw[i] * dphi;
      for (int l = 0; l <= lmax; l++)
        for (int m = 0; m <= l; m++) {
          int k = l * (l + 1) / 2 + m;
          double fm = (m == 0) ? 1.0 : 2.0;
          a_lm[k] += fm * w * r * P[k] * cos(m * phi);
          b_lm[k] += fm * w * r * P[k] * sin(m * phi);
        }
    }
  }
  free(gx); free(gw); free(P);
  return 0;
}
