// pair_kernel.cuh — the SPHERHARM pair kernel (kernels K3-K5), sm_100a.
//
// Replaces Pair::compute of `pair_style spherharm` (reference source NOT IN MOUNT; the algorithm
// is the one BASELINE.json:5 describes, made precise in SURVEY Appendix A.4/A.5 and DESIGN §3).
//
// One persistent CTA per SM slot pulls unordered pairs (i,j) from an atomic work counter.  For
// each of the two directions (nodes of a tested against the surface of b):
//   1. relative pose M = Rb^T Ra, t = Rb^T (c_a - c_b) computed once per pair into shared memory;
//      b's folded coefficient table staged in shared memory (skipped when already resident);
//   2. CULL   : every warp walks 32-node chunks of a's node table (coalesced SoA loads), transforms
//               them into b's frame, ballots the bounding-sphere survivors and compacts their node
//               indices (uint16) into one CTA-wide shared list; nodes inside b's inscribed sphere
//               set their bit of the inside-mask directly;
//   3. EVAL   : full warps over the compacted list evaluate r_b with the folded recurrences
//               (1 DMUL + 3 DFMA per (l,m) term) and set the inside bit when rho < r_b;
//   4. SUM    : a fixed-order pass over the inside-mask accumulates S, A, T, G in a's frame;
//               warp-shuffle + shared-memory reduction in a fixed order (deterministic).
// Thread 0 then applies the contact law and writes the per-pair record and the two per-entry
// force/torque slots that the gather kernel sums in a fixed order (no floating-point atomics).
#pragma once
#include "device_math.cuh"

namespace shgpu {

#define SH_MAX_SHAPES 64

struct PairArgs {
  const DevShape *shapes;
  const double *c, *Rs, *x;   // SoA with stride: c[d*stride+i], Rs[e*stride+i], x[d*stride+i]
  const int *shape;
  int stride;
  const int *pair_i, *pair_j, *pair_eij, *pair_eji;
  const int *pair_img;        // periodic image of the pair, 2 bits per dimension (n + 1), n = rint((c_i - c_j) / L) at the list build
  int npairs;
  double *slot;               // 6 x slot_stride : F(3), torque(3) per neighbor-list entry
  int slot_stride;
  double *pres;               // 14 x pres_stride: V,E,F3,ti3,tj3,xc3
  int pres_stride;
  const double *pk, *pm;      // SH_MAX_SHAPES^2 stiffness / exponent
  double boxlen[3];
  int periodic[3];
  int *work_counter;
  unsigned long long *counters;  // [0]=pairs [1]=nodes transformed [2]=evaluated [3]=inside [4]=pairs with a ghost
  int max_terms, max_nq;      // shared-memory sizing
  int nlocal;                 // atoms >= nlocal are ghosts (counters[4] counts pairs with a ghost)
  const int *pair_list;       // optional indirection: process pairs pair_list[0..npairs) (deep-contact list)
  const int *npairs_dev;      // optional: number of pairs lives on the device (deep-contact list of the split pipeline)
  // dissipative contact terms (oracle A.5b): viscous normal damping + Coulomb-capped tangential friction
  int dissip;
  const double *v, *L, *q;    // SoA velocities, angular momenta, quaternions (ghosts carry theirs when dissip is on)
  const double *pgn, *pgt, *pmu;
};

// minimum-image separation d = c_i - c_j - L n with the image n stored at the neighbor build: the same value, bit for
// bit, as the oracle's d - L rint(d / L) (n = 0 leaves d untouched), without three FP64 divisions per pair and kernel
__device__ __forceinline__ void pair_separation(const PairArgs &A, int p, int i, int j, double d[3]) {
  const int st = A.stride, img = A.pair_img[p];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double dk = A.c[k * st + i] - A.c[k * st + j];
    const int n = ((img >> (2 * k)) & 3) - 1;
    if (n != 0) dk = dk - A.boxlen[k] * (double)n;
    d[k] = dk;
  }
}

// the CPU checker's contact_dissipation (spec A.5b), same operations in the same order.  out: [2..4] F on i, [5..7] tau_i,
// [8..10] tau_j, [11..13] overlap centroid (i's periodic image); d = c_i - c_j (minimum image); lj = c_j - x_j.
__device__ __forceinline__ void contact_dissipation(const PairArgs &A, int i, int j, int shp_i, int shp_j, const double d[3],
                                                    const double lj[3], double out[14]) {
  const double gn = A.pgn[shp_i * SH_MAX_SHAPES + shp_j], gt = A.pgt[shp_i * SH_MAX_SHAPES + shp_j], mu = A.pmu[shp_i * SH_MAX_SHAPES + shp_j];
  if (!(gn > 0.0 || (gt > 0.0 && mu > 0.0))) return;
  const int st = A.stride;
  double *F = out + 2, *ti = out + 5, *tj = out + 8;
  const double *xc = out + 11;
  const double fn2 = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
  if (!(fn2 > 0.0)) return;
  const double fn = sqrt(fn2);
  const double nh[3] = {F[0] / fn, F[1] / fn, F[2] / fn};
  double xi[3], xj[3], vi[3], vj[3], Li[3], Lj[3], qi[4], qj[4], wi[3], wj[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    xi[r] = A.x[r * st + i];
    xj[r] = (A.c[r * st + i] - d[r]) - lj[r];
    vi[r] = A.v[r * st + i]; vj[r] = A.v[r * st + j];
    Li[r] = A.L[r * st + i]; Lj[r] = A.L[r * st + j];
  }
#pragma unroll
  for (int r = 0; r < 4; r++) { qi[r] = A.q[r * st + i]; qj[r] = A.q[r * st + j]; }
  omega_from_angmom(qi, Li, A.shapes[shp_i].inertia, wi);
  omega_from_angmom(qj, Lj, A.shapes[shp_j].inertia, wj);
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
#pragma unroll
  for (int r = 0; r < 3; r++) F[r] += Fd[r];
  ti[0] += ri[1] * Fd[2] - ri[2] * Fd[1]; ti[1] += ri[2] * Fd[0] - ri[0] * Fd[2]; ti[2] += ri[0] * Fd[1] - ri[1] * Fd[0];
  tj[0] -= rj[1] * Fd[2] - rj[2] * Fd[1]; tj[1] -= rj[2] * Fd[0] - rj[0] * Fd[2]; tj[2] -= rj[0] * Fd[1] - rj[1] * Fd[0];
}

__host__ __device__ inline size_t pair_smem_bytes(int max_terms, int max_nq, int nwarps) {
  size_t b = 0;
  b += (size_t)max_terms * sizeof(double2);          // ab
  b += (size_t)max_terms * sizeof(double);           // Ap
  b += (size_t)(15 + 20 + nwarps * 10) * sizeof(double);  // pose, dir sums, reduction scratch
  b += (size_t)((max_nq + 31) / 32) * sizeof(unsigned);   // inside mask
  b += (size_t)((max_nq + 31) / 32) * 32 * sizeof(unsigned short);  // survivor list
  b += 16 * sizeof(int);
  return (b + 15) & ~(size_t)15;
}

template <int NT>
__global__ void __launch_bounds__(NT) pair_kernel(PairArgs A) {
  constexpr int W = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *s_ab = reinterpret_cast<double2 *>(smem_raw);
  double *s_Ap = reinterpret_cast<double *>(s_ab + A.max_terms);
  double *s_pose = s_Ap + A.max_terms;     // M[9], t[3], x0[3]
  double *s_dir = s_pose + 15;             // [2][10]
  double *s_red = s_dir + 20;              // [W][10]
  unsigned *s_mask = reinterpret_cast<unsigned *>(s_red + W * 10);
  const int max_chunks = (A.max_nq + 31) / 32;
  unsigned short *s_list = reinterpret_cast<unsigned short *>(s_mask + max_chunks);
  int *s_int = reinterpret_cast<int *>(s_list + max_chunks * 32);  // [0]=pair [1]=nsurv [2]=loaded shape

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long n_eval = 0, n_inside = 0, n_trans = 0, n_pairs = 0, n_gh = 0;  // thread 0 only
  if (tid == 0) s_int[2] = -1;

  for (;;) {
    if (tid == 0) s_int[0] = atomicAdd(A.work_counter, 1);
    __syncthreads();
    if (s_int[0] >= A.npairs) break;
    const int p = A.pair_list ? A.pair_list[s_int[0]] : s_int[0];
    const int i = A.pair_i[p], j = A.pair_j[p];
    const int st = A.stride;
    // minimum-image separation d = c_i - c_j (same ops as the oracle)
    double d[3];
    pair_separation(A, p, i, j, d);
    const int shp_i = A.shape[i], shp_j = A.shape[j];

    for (int dir = 0; dir < 2; dir++) {
      const int a = dir ? j : i, b = dir ? i : j;
      const int sa_id = dir ? shp_j : shp_i, sb_id = dir ? shp_i : shp_j;
      const DevShape &sa = A.shapes[sa_id];
      const DevShape &sb = A.shapes[sb_id];
      const double sgn = dir ? -1.0 : 1.0;
      // ---- relative pose: lanes 0..14 of warp 0 each produce one element
      if (tid < 15) {
        const double dd0 = sgn * d[0], dd1 = sgn * d[1], dd2 = sgn * d[2];
        if (tid < 9) {
          const int r = tid / 3, k = tid - 3 * r;
          double m = A.Rs[(0 + r) * st + b] * A.Rs[(0 + k) * st + a];
          m = fma(A.Rs[(3 + r) * st + b], A.Rs[(3 + k) * st + a], m);
          m = fma(A.Rs[(6 + r) * st + b], A.Rs[(6 + k) * st + a], m);
          s_pose[tid] = m;
        } else if (tid < 12) {
          const int r = tid - 9;
          double tt = A.Rs[(0 + r) * st + b] * dd0;
          tt = fma(A.Rs[(3 + r) * st + b], dd1, tt);
          tt = fma(A.Rs[(6 + r) * st + b], dd2, tt);
          s_pose[tid] = tt;
        } else {
          const int r = tid - 12;
          double hx = A.Rs[(0 + r) * st + a] * dd0;
          hx = fma(A.Rs[(3 + r) * st + a], dd1, hx);
          hx = fma(A.Rs[(6 + r) * st + a], dd2, hx);
          s_pose[tid] = -0.5 * hx;
        }
      }
      if (tid == 0) s_int[1] = 0;
      // ---- stage b's folded tables (skip when resident)
      if (s_int[2] != sb_id) {
        const int T = sb.nterms;
        for (int t = tid; t < T; t += NT) { s_ab[t] = sb.ab[t]; s_Ap[t] = sb.Ap[t]; }
      }
      __syncthreads();
      if (tid == 0) s_int[2] = sb_id;
      const double M00 = s_pose[0], M01 = s_pose[1], M02 = s_pose[2];
      const double M10 = s_pose[3], M11 = s_pose[4], M12 = s_pose[5];
      const double M20 = s_pose[6], M21 = s_pose[7], M22 = s_pose[8];
      const double t0 = s_pose[9], t1 = s_pose[10], t2 = s_pose[11];
      const double rmax2 = sb.rmax2, rmin2 = sb.rmin2;
      const int nq = sa.nq, nchunks = sa.nchunks;
      const double *__restrict__ px = sa.px, *__restrict__ py = sa.py, *__restrict__ pz = sa.pz;

      // ---- CULL
      for (int ch = warp; ch < nchunks; ch += W) {
        const int k = ch * 32 + lane;
        bool surv = false, autoin = false;
        if (k < nq) {
          const double p0 = px[k], p1 = py[k], p2 = pz[k];
          double s0 = fma(M00, p0, t0); s0 = fma(M01, p1, s0); s0 = fma(M02, p2, s0);
          double s1 = fma(M10, p0, t1); s1 = fma(M11, p1, s1); s1 = fma(M12, p2, s1);
          double s2 = fma(M20, p0, t2); s2 = fma(M21, p1, s2); s2 = fma(M22, p2, s2);
          const double rho2 = fma(s2, s2, fma(s1, s1, s0 * s0));
          if (rho2 < rmax2) { if (rho2 <= rmin2) autoin = true; else surv = true; }
        }
        const unsigned am = __ballot_sync(0xffffffffu, autoin);
        const unsigned sm = __ballot_sync(0xffffffffu, surv);
        int basepos = 0;
        if (lane == 0) { s_mask[ch] = am; if (sm) basepos = atomicAdd(&s_int[1], __popc(sm)); }
        basepos = __shfl_sync(0xffffffffu, basepos, 0);
        if (surv) s_list[basepos + __popc(sm & ((1u << lane) - 1u))] = (unsigned short)k;
      }
      __syncthreads();

      // ---- EVAL (full warps over the compacted list)
      const int nsurv = s_int[1];
      const int L = sb.lmax;
      for (int g = tid; g < nsurv; g += NT) {
        const int k = s_list[g];
        const double p0 = px[k], p1 = py[k], p2 = pz[k];
        double s0 = fma(M00, p0, t0); s0 = fma(M01, p1, s0); s0 = fma(M02, p2, s0);
        double s1 = fma(M10, p0, t1); s1 = fma(M11, p1, s1); s1 = fma(M12, p2, s1);
        double s2 = fma(M20, p0, t2); s2 = fma(M21, p1, s2); s2 = fma(M22, p2, s2);
        const double rho2 = fma(s2, s2, fma(s1, s1, s0 * s0));
        double rho;
        const double r = sh_radius_folded(L, s_Ap, s_ab, s0, s1, s2, rho2, rho);
        if (rho < r) atomicOr(&s_mask[k >> 5], 1u << (k & 31));
      }
      __syncthreads();

      // ---- SUM over inside nodes, fixed order
      const double x00 = s_pose[12], x01 = s_pose[13], x02 = s_pose[14];
      double S0 = 0, S1 = 0, S2 = 0, Av = 0, T0 = 0, T1 = 0, T2 = 0, G0 = 0, G1 = 0, G2 = 0;
      int cnt = 0;
      for (int ch = warp; ch < nchunks; ch += W) {
        const unsigned word = s_mask[ch];
        if ((word >> lane) & 1u) {
          const int k = ch * 32 + lane;
          const double p0 = px[k], p1 = py[k], p2 = pz[k];
          const double n0 = sa.nx[k], n1 = sa.ny[k], n2 = sa.nz[k];
          const double dp0 = p0 - x00, dp1 = p1 - x01, dp2 = p2 - x02;
          const double dn = fma(dp2, n2, fma(dp1, n1, dp0 * n0));
          S0 += n0; S1 += n1; S2 += n2;
          Av += dn;
          T0 += fma(p1, n2, -(p2 * n1));
          T1 += fma(p2, n0, -(p0 * n2));
          T2 += fma(p0, n1, -(p1 * n0));
          G0 = fma(dp0, dn, G0); G1 = fma(dp1, dn, G1); G2 = fma(dp2, dn, G2);
          cnt++;
        }
      }
      S0 = warp_sum(S0); S1 = warp_sum(S1); S2 = warp_sum(S2); Av = warp_sum(Av);
      T0 = warp_sum(T0); T1 = warp_sum(T1); T2 = warp_sum(T2);
      G0 = warp_sum(G0); G1 = warp_sum(G1); G2 = warp_sum(G2);
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0) {
        double *rr = s_red + warp * 10;
        rr[0] = S0; rr[1] = S1; rr[2] = S2; rr[3] = Av; rr[4] = T0; rr[5] = T1; rr[6] = T2;
        rr[7] = G0; rr[8] = G1; rr[9] = G2;
        s_int[4 + warp] = cnt;
      }
      __syncthreads();
      if (tid < 10) {
        double acc = 0;
        for (int w = 0; w < W; w++) acc += s_red[w * 10 + tid];
        s_dir[dir * 10 + tid] = acc;
      }
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < W; w++) tot += s_int[4 + w];
        s_int[2 + 12 + dir] = tot;  // s_int[14], s_int[15]
        n_eval += nsurv; n_inside += tot; n_trans += nq;
      }
      __syncthreads();
    }  // dir

    // ---- contact law + outputs
    if (tid == 0) {
      n_pairs++; n_gh += (j >= A.nlocal);
      const double *Ri_ = A.Rs, *dij = s_dir, *dji = s_dir + 10;
      // rotate body-frame sums to the space frame
      double Sij[3], Tij[3], Gij[3], Sji[3], Tji[3], Gji[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double a0 = Ri_[(3 * r) * st + i], a1 = Ri_[(3 * r + 1) * st + i], a2 = Ri_[(3 * r + 2) * st + i];
        Sij[r] = a0 * dij[0] + a1 * dij[1] + a2 * dij[2];
        Tij[r] = a0 * dij[4] + a1 * dij[5] + a2 * dij[6];
        Gij[r] = 0.25 * (a0 * dij[7] + a1 * dij[8] + a2 * dij[9]);
        const double b0 = Ri_[(3 * r) * st + j], b1 = Ri_[(3 * r + 1) * st + j], b2 = Ri_[(3 * r + 2) * st + j];
        Sji[r] = b0 * dji[0] + b1 * dji[1] + b2 * dji[2];
        Tji[r] = b0 * dji[4] + b1 * dji[5] + b2 * dji[6];
        Gji[r] = 0.25 * (b0 * dji[7] + b1 * dji[8] + b2 * dji[9]);
      }
      const double V = dij[3] / 3.0 + dji[3] / 3.0;
      const int ninside = s_int[14] + s_int[15];
      double out[14];
#pragma unroll
      for (int r = 0; r < 14; r++) out[r] = 0.0;
      if (ninside > 0 && V > 0) {
        const double kk = A.pk[shp_i * SH_MAX_SHAPES + shp_j], mm = A.pm[shp_i * SH_MAX_SHAPES + shp_j];
        double E, pr;
        if (mm == 1.0) { E = kk * V; pr = kk; }
        else { const double pw = pow(V, mm - 1.0); E = kk * pw * V; pr = mm * kk * pw; }
        out[0] = V; out[1] = E;
        double li[3], lj[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          li[r] = A.c[r * st + i] - A.x[r * st + i];
          lj[r] = A.c[r * st + j] - A.x[r * st + j];
        }
        const double Ti[3] = {Tij[0] + (li[1] * Sij[2] - li[2] * Sij[1]), Tij[1] + (li[2] * Sij[0] - li[0] * Sij[2]),
                              Tij[2] + (li[0] * Sij[1] - li[1] * Sij[0])};
        const double Tj[3] = {Tji[0] + (lj[1] * Sji[2] - lj[2] * Sji[1]), Tji[1] + (lj[2] * Sji[0] - lj[0] * Sji[2]),
                              Tji[2] + (lj[0] * Sji[1] - lj[1] * Sji[0])};
#pragma unroll
        for (int r = 0; r < 3; r++) {
          out[2 + r] = -pr * (0.5 * (Sij[r] - Sji[r]));
          out[5 + r] = -pr * Ti[r];
          out[8 + r] = -pr * Tj[r];
          out[11 + r] = (A.c[r * st + i] - 0.5 * d[r]) + (Gij[r] + Gji[r]) / V;
        }
        if (A.dissip) contact_dissipation(A, i, j, shp_i, shp_j, d, lj, out);
      }
#pragma unroll
      for (int r = 0; r < 14; r++) A.pres[(size_t)r * A.pres_stride + p] = out[r];
      const int eij = A.pair_eij[p], eji = A.pair_eji[p];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        A.slot[(size_t)r * A.slot_stride + eij] = out[2 + r];
        A.slot[(size_t)(3 + r) * A.slot_stride + eij] = out[5 + r];
        if (eji >= 0) {
          A.slot[(size_t)r * A.slot_stride + eji] = -out[2 + r];
          A.slot[(size_t)(3 + r) * A.slot_stride + eji] = out[8 + r];
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    atomicAdd(&A.counters[0], n_pairs);
    atomicAdd(&A.counters[1], n_trans);
    atomicAdd(&A.counters[2], n_eval);
    atomicAdd(&A.counters[3], n_inside);
    atomicAdd(&A.counters[4], n_gh);
  }
}

}  // namespace shgpu
