// pair_warp_kernel.cuh — warp-per-pair SPHERHARM pair kernel (default), sm_100a.
//
// Same arithmetic contract as pair_kernel.cuh; different schedule:
//   * every warp is persistent and pulls unordered pairs from an atomic counter: no block barriers,
//     no idle warps while a small contact is being processed;
//   * all shapes' folded coefficient tables are resident in shared memory (one copy per CTA);
//   * WINDOW: instead of transforming all N_theta x N_phi nodes of a, only the rows/columns of a's
//     node grid that can geometrically reach b's bounding sphere are visited.  The window is a
//     conservative FP32 bound (cone around the centre line, with margins); it only ever rejects
//     nodes whose exact test would fail, so the exact FP64 test below still decides every node and
//     the result is identical to a full scan (the parity tests check the evaluated/inside counters
//     against the oracle's full scan, bit for bit);
//   * DIRECTION-CELL BOUND: a node that passes the bounding-sphere test is first compared with a
//     conservative upper bound of r_b over its cube-map direction cell (6 x 24 x 24 cells per shape, built
//     at shape load); rho^2 >= bound^2 proves "outside" without evaluating the series.  In a jammed packing
//     this removes ~90 % of the SH evaluations; the decisions are unchanged (inside counters stay
//     bit-equal to the oracle, which evaluates every survivor);
//   * STREAMING COMPACTION: survivors of the exact bounding-sphere test go to a 128-entry per-warp
//     ring; whenever >= 64 are queued the warp evaluates r_b for two points per lane with the folded
//     recurrences (1 DMUL + 3 DFMA per (l,m) term and point; the two points share every coefficient
//     load and form independent dependency chains) — every lane busy except in the last flush;
//   * inside nodes are accumulated at once; the visiting order is a pure function of the pair, so the
//     floating-point sums are bitwise reproducible run to run (no floating-point atomics).
#pragma once
#include "pair_kernel.cuh"

namespace shgpu {

struct DirAcc {
  double S0, S1, S2, A, T0, T1, T2, G0, G1, G2;
  int cnt;
};

template <int NW, bool SMEM_TABLES>
__global__ void __launch_bounds__(NW * 32) pair_warp_kernel(PairArgs A, int nshapes, int total_terms, int use_bounds) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *s_ab = reinterpret_cast<double2 *>(smem_raw);
  double *s_Ap = reinterpret_cast<double *>(s_ab + (SMEM_TABLES ? total_terms : 0));
  double *s_dir = s_Ap + (SMEM_TABLES ? total_terms : 0);            // [NW][10] direction-0 sums
  double *s_pose = s_dir + NW * 10;                                  // [NW][16] M[9], t[3], x0[3]
  unsigned short *s_ring = reinterpret_cast<unsigned short *>(s_pose + NW * 16);  // [NW][128]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // deep-contact list of the split pipeline: its length lives on the device (no host round trip)
  const int npairs = A.npairs_dev ? *A.npairs_dev : A.npairs;
  if (npairs == 0) return;
  if (SMEM_TABLES) {
    for (int s = 0; s < nshapes; s++) {
      const DevShape &sh = A.shapes[s];
      for (int t = tid; t < sh.nterms4 + 4; t += NW * 32) { s_ab[sh.tab_off + t] = sh.ab[t]; s_Ap[sh.tab_off + t] = sh.Ap[t]; }
    }
    __syncthreads();
  }
  unsigned short *ring = s_ring + warp * 128;
  double *dir0 = s_dir + warp * 10;
  double *pose = s_pose + warp * 16;
  unsigned long long n_eval = 0, n_inside = 0, n_trans = 0, n_pairs = 0, n_gh = 0;
  const int st = A.stride;

  for (;;) {
    int p = 0;
    if (lane == 0) p = atomicAdd(A.work_counter, 1);
    p = __shfl_sync(0xffffffffu, p, 0);
    if (p >= npairs) break;
    if (A.pair_list) p = A.pair_list[p];
    const int i = A.pair_i[p], j = A.pair_j[p];
    double d[3];
    pair_separation(A, p, i, j, d);
    const int shp_i = A.shape[i], shp_j = A.shape[j];
    DirAcc acc;
    int ninside_pair = 0;

    for (int dir = 0; dir < 2; dir++) {
      const int a = dir ? j : i, b = dir ? i : j;
      const DevShape &sa = A.shapes[dir ? shp_j : shp_i];
      const DevShape &sb = A.shapes[dir ? shp_i : shp_j];
      const double sgn = dir ? -1.0 : 1.0;
      const double dd0 = sgn * d[0], dd1 = sgn * d[1], dd2 = sgn * d[2];
      // relative pose: lanes 0..14 each produce one element (same fma chains as the oracle) into
      // this warp's shared scratch; kept there (not in registers) to leave room for the eval pipeline
      __syncwarp();
      if (lane < 15) {
        double val;
        if (lane < 9) {
          const int r = lane / 3, k = lane - 3 * r;
          double m = A.Rs[(0 + r) * st + b] * A.Rs[(0 + k) * st + a];
          m = fma(A.Rs[(3 + r) * st + b], A.Rs[(3 + k) * st + a], m);
          m = fma(A.Rs[(6 + r) * st + b], A.Rs[(6 + k) * st + a], m);
          val = m;
        } else if (lane < 12) {
          const int r = lane - 9;
          double tt = A.Rs[(0 + r) * st + b] * dd0;
          tt = fma(A.Rs[(3 + r) * st + b], dd1, tt);
          tt = fma(A.Rs[(6 + r) * st + b], dd2, tt);
          val = tt;
        } else {
          const int r = lane - 12;
          double hx = A.Rs[(0 + r) * st + a] * dd0;
          hx = fma(A.Rs[(3 + r) * st + a], dd1, hx);
          hx = fma(A.Rs[(6 + r) * st + a], dd2, hx);
          val = -0.5 * hx;
        }
        pose[lane] = val;
      }
      __syncwarp();
      const double *M = pose, *t = pose + 9, *x0 = pose + 12;
      acc.S0 = acc.S1 = acc.S2 = acc.A = acc.T0 = acc.T1 = acc.T2 = acc.G0 = acc.G1 = acc.G2 = 0.0;
      acc.cnt = 0;
      const double rmax2 = sb.rmax2, rmin2 = sb.rmin2;
      const double *__restrict__ nodes = sa.px;   // SoA block: px,py,pz,nx,ny,nz each nq long
      const int nq = sa.nq;
      const double *tabAp = SMEM_TABLES ? (s_Ap + sb.tab_off) : sb.Ap;
      const double2 *tabab = SMEM_TABLES ? (s_ab + sb.tab_off) : sb.ab;
      const int L = sb.lmax;
      const float2 *__restrict__ cube = sb.cube_ul;
      const int cn = sb.cube_n;
      int queued = 0;  // ring holds entries [0, queued)

      auto accumulate = [&](int k, double p0, double p1, double p2) {
        const double n0 = nodes[3 * nq + k], n1 = nodes[4 * nq + k], n2 = nodes[5 * nq + k];
        const double dp0 = p0 - x0[0], dp1 = p1 - x0[1], dp2 = p2 - x0[2];
        const double dn = fma(dp2, n2, fma(dp1, n1, dp0 * n0));
        acc.S0 += n0; acc.S1 += n1; acc.S2 += n2;
        acc.A += dn;
        acc.T0 += fma(p1, n2, -(p2 * n1));
        acc.T1 += fma(p2, n0, -(p0 * n2));
        acc.T2 += fma(p0, n1, -(p1 * n0));
        acc.G0 = fma(dp0, dn, acc.G0); acc.G1 = fma(dp1, dn, acc.G1); acc.G2 = fma(dp2, dn, acc.G2);
        acc.cnt++;
      };
      auto evaluate = [&](int count) {  // evaluate ring[0..count) with lanes < count
        if (lane < count) {
          const int k = ring[lane];
          const double p0 = nodes[k], p1 = nodes[nq + k], p2 = nodes[2 * nq + k];
          double s0 = fma(M[0], p0, t[0]); s0 = fma(M[1], p1, s0); s0 = fma(M[2], p2, s0);
          double s1 = fma(M[3], p0, t[1]); s1 = fma(M[4], p1, s1); s1 = fma(M[5], p2, s1);
          double s2 = fma(M[6], p0, t[2]); s2 = fma(M[7], p1, s2); s2 = fma(M[8], p2, s2);
          const double rho2 = fma(s2, s2, fma(s1, s1, s0 * s0));
          double rho;
          const double r = sh_radius_folded(L, tabAp, tabab, s0, s1, s2, rho2, rho);
          if (rho < r) accumulate(k, p0, p1, p2);
        }
      };
      // two points per lane: ring[lane] and ring[32+lane]; slot B is valid for lane < count-32
      auto evaluate2 = [&](int count) {
        const int kA = ring[lane];
        const bool validB = lane < count - 32;
        const int kB = validB ? ring[32 + lane] : kA;
        double sA[3], sB[3];
        {
          const double p0 = nodes[kA], p1 = nodes[nq + kA], p2 = nodes[2 * nq + kA];
          double s0 = fma(M[0], p0, t[0]); s0 = fma(M[1], p1, s0); s0 = fma(M[2], p2, s0);
          double s1 = fma(M[3], p0, t[1]); s1 = fma(M[4], p1, s1); s1 = fma(M[5], p2, s1);
          double s2 = fma(M[6], p0, t[2]); s2 = fma(M[7], p1, s2); s2 = fma(M[8], p2, s2);
          sA[0] = s0; sA[1] = s1; sA[2] = s2;
        }
        {
          const double p0 = nodes[kB], p1 = nodes[nq + kB], p2 = nodes[2 * nq + kB];
          double s0 = fma(M[0], p0, t[0]); s0 = fma(M[1], p1, s0); s0 = fma(M[2], p2, s0);
          double s1 = fma(M[3], p0, t[1]); s1 = fma(M[4], p1, s1); s1 = fma(M[5], p2, s1);
          double s2 = fma(M[6], p0, t[2]); s2 = fma(M[7], p1, s2); s2 = fma(M[8], p2, s2);
          sB[0] = s0; sB[1] = s1; sB[2] = s2;
        }
        const double rhoA2 = fma(sA[2], sA[2], fma(sA[1], sA[1], sA[0] * sA[0]));
        const double rhoB2 = fma(sB[2], sB[2], fma(sB[1], sB[1], sB[0] * sB[0]));
        double rhoA, rhoB, rA, rB;
        sh_radius_folded_x2(L, tabAp, tabab, sA, rhoA2, sB, rhoB2, rhoA, rhoB, rA, rB);
        if (rhoA < rA) accumulate(kA, nodes[kA], nodes[nq + kA], nodes[2 * nq + kA]);
        if (validB && rhoB < rB) accumulate(kB, nodes[kB], nodes[nq + kB], nodes[2 * nq + kB]);
      };

      // ---- conservative window on a's node grid (FP32 with margins; see header comment)
      const double e0 = 2.0 * x0[0], e1 = 2.0 * x0[1], e2 = 2.0 * x0[2];   // b's origin in a's frame
      const double D2 = e0 * e0 + e1 * e1 + e2 * e2, D = sqrt(D2);
      bool skip = D >= (sa.rmax + sb.rmax) * (1.0 + 1e-9);
      float cosA = -2.0f, xe = 1.0f, se = 0.0f, phie = 0.0f;
      if (!skip && D > 1e-9 * (sa.rmax + sb.rmax)) {
        const double q = D2 - rmax2;
        double rc = sa.rmin;
        if (q > 0) rc = fmin(fmax(sqrt(q), sa.rmin), sa.rmax);
        const double g = (rc * rc + q) / (2.0 * rc * D);
        cosA = (float)g - 3e-5f;
        if (cosA >= 1.0f) skip = true;
        xe = (float)(e2 / D);
        xe = fminf(1.0f, fmaxf(-1.0f, xe));
        se = sqrtf(fmaxf(0.0f, 1.0f - xe * xe));
        phie = atan2f((float)e1, (float)e0);
        if (phie < 0.0f) phie += 6.2831853f;
      }
      if (!skip) {
        const int nth = sa.n_theta, nph = sa.n_phi;
        const float inv_dphi = (float)nph * 0.15915494f;
        for (int rb = 0; rb < nth; rb += 32) {
          // each lane classifies one row of this batch
          const int row = rb + lane;
          int c0 = 0, ccount = 0;
          if (row < nth) {
            if (cosA <= -1.0f) { ccount = nph; }
            else {
              const float xa = sa.row_x[row];
              const float sarow = sqrtf(fmaxf(0.0f, 1.0f - xa * xa));
              const float ss = sarow * se, xx = xa * xe;
              if (xx + ss >= cosA) {
                float cd = (ss > 1e-12f) ? (cosA - xx) / ss : -2.0f;
                if (cd <= -1.0f) ccount = nph;
                else {
                  const float dl = acosf(fminf(cd, 1.0f)) + 2e-4f;
                  const int b0 = (int)ceilf((phie - dl) * inv_dphi - 0.5f);
                  const int b1 = (int)floorf((phie + dl) * inv_dphi - 0.5f);
                  ccount = b1 - b0 + 1;
                  if (ccount >= nph) { ccount = nph; c0 = 0; }
                  else if (ccount > 0) { c0 = b0 % nph; if (c0 < 0) c0 += nph; }
                  else ccount = 0;
                }
              }
            }
          }
          unsigned rows = __ballot_sync(0xffffffffu, ccount > 0);
          while (rows) {
            const int rl = __ffs(rows) - 1;
            rows &= rows - 1;
            const int rc0 = __shfl_sync(0xffffffffu, c0, rl), rcount = __shfl_sync(0xffffffffu, ccount, rl);
            const int rowbase = (rb + rl) * nph;
            for (int cb = 0; cb < rcount; cb += 32) {
              const int cc = cb + lane;
              bool surv = false;
              int k = 0;
              if (cc < rcount) {
                int col = rc0 + cc;
                if (col >= nph) col -= nph;
                k = rowbase + col;
                const double p0 = nodes[k], p1 = nodes[nq + k], p2 = nodes[2 * nq + k];
                double s0 = fma(M[0], p0, t[0]); s0 = fma(M[1], p1, s0); s0 = fma(M[2], p2, s0);
                double s1 = fma(M[3], p0, t[1]); s1 = fma(M[4], p1, s1); s1 = fma(M[5], p2, s1);
                double s2 = fma(M[6], p0, t[2]); s2 = fma(M[7], p1, s2); s2 = fma(M[8], p2, s2);
                const double rho2 = fma(s2, s2, fma(s1, s1, s0 * s0));
                if (rho2 < rmax2) {
                  if (rho2 <= rmin2) accumulate(k, p0, p1, p2);   // inside b's inscribed sphere
                  else if (use_bounds) {
                    // direction cell of the cube map (FP32 ratios; the table is conservative across borders)
                    const float fx = (float)s0, fy = (float)s1, fz = (float)s2;
                    const float ax = fabsf(fx), ay = fabsf(fy), az = fabsf(fz);
                    int face; float ma, uu, vv;
                    if (ax >= ay && ax >= az) { face = fx > 0 ? 0 : 1; ma = ax; uu = fy; vv = fz; }
                    else if (ay >= az) { face = fy > 0 ? 2 : 3; ma = ay; uu = fx; vv = fz; }
                    else { face = fz > 0 ? 4 : 5; ma = az; uu = fx; vv = fy; }
                    const float im = 1.0f / fmaxf(ma, 1e-30f), hn = 0.5f * (float)cn;
                    const int iu = min(cn - 1, max(0, (int)((uu * im + 1.0f) * hn)));
                    const int iv = min(cn - 1, max(0, (int)((vv * im + 1.0f) * hn)));
                    // proven bounds of r^2 over the cell: below the lower bound the node is inside, at or above
                    // the upper bound it is outside, in between the series decides
                    const float2 ul = __ldg(&cube[(face * cn + iu) * cn + iv]);
                    if (rho2 <= (double)ul.y) accumulate(k, p0, p1, p2);
                    else surv = rho2 < (double)ul.x;
                  } else surv = true;
                }
              }
              n_trans += min(32, rcount - cb);
              const unsigned sm = __ballot_sync(0xffffffffu, surv);
              if (sm) {
                if (surv) ring[queued + __popc(sm & ((1u << lane) - 1u))] = (unsigned short)k;
                queued += __popc(sm);
                __syncwarp();
                if (queued >= 64) {
                  evaluate2(64);
                  n_eval += 64;
                  __syncwarp();
                  // move the remainder (< 32 entries) to the front
                  const int rem = queued - 64;
                  unsigned short mv = 0;
                  if (lane < rem) mv = ring[64 + lane];
                  __syncwarp();
                  if (lane < rem) ring[lane] = mv;
                  queued = rem;
                  __syncwarp();
                }
              }
            }
          }
        }
        if (queued > 32) { evaluate2(queued); n_eval += queued; __syncwarp(); }
        else if (queued > 0) { evaluate(queued); n_eval += queued; __syncwarp(); }
      }
      // ---- warp reduction (fixed butterfly order)
      acc.S0 = warp_sum(acc.S0); acc.S1 = warp_sum(acc.S1); acc.S2 = warp_sum(acc.S2); acc.A = warp_sum(acc.A);
      acc.T0 = warp_sum(acc.T0); acc.T1 = warp_sum(acc.T1); acc.T2 = warp_sum(acc.T2);
      acc.G0 = warp_sum(acc.G0); acc.G1 = warp_sum(acc.G1); acc.G2 = warp_sum(acc.G2);
      acc.cnt = __reduce_add_sync(0xffffffffu, acc.cnt);
      ninside_pair += acc.cnt;
      if (dir == 0) {
        if (lane == 0) {
          dir0[0] = acc.S0; dir0[1] = acc.S1; dir0[2] = acc.S2; dir0[3] = acc.A; dir0[4] = acc.T0; dir0[5] = acc.T1;
          dir0[6] = acc.T2; dir0[7] = acc.G0; dir0[8] = acc.G1; dir0[9] = acc.G2;
        }
        __syncwarp();
      }
    }  // dir

    // ---- contact law + outputs; uniform across the warp, lane 0 stores
    {
      n_pairs++; n_inside += ninside_pair; n_gh += (j >= A.nlocal);
      const double dij[10] = {dir0[0], dir0[1], dir0[2], dir0[3], dir0[4], dir0[5], dir0[6], dir0[7], dir0[8], dir0[9]};
      const double dji[10] = {acc.S0, acc.S1, acc.S2, acc.A, acc.T0, acc.T1, acc.T2, acc.G0, acc.G1, acc.G2};
      double Sij[3], Tij[3], Gij[3], Sji[3], Tji[3], Gji[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double a0 = A.Rs[(3 * r) * st + i], a1 = A.Rs[(3 * r + 1) * st + i], a2 = A.Rs[(3 * r + 2) * st + i];
        Sij[r] = a0 * dij[0] + a1 * dij[1] + a2 * dij[2];
        Tij[r] = a0 * dij[4] + a1 * dij[5] + a2 * dij[6];
        Gij[r] = 0.25 * (a0 * dij[7] + a1 * dij[8] + a2 * dij[9]);
        const double b0 = A.Rs[(3 * r) * st + j], b1 = A.Rs[(3 * r + 1) * st + j], b2 = A.Rs[(3 * r + 2) * st + j];
        Sji[r] = b0 * dji[0] + b1 * dji[1] + b2 * dji[2];
        Tji[r] = b0 * dji[4] + b1 * dji[5] + b2 * dji[6];
        Gji[r] = 0.25 * (b0 * dji[7] + b1 * dji[8] + b2 * dji[9]);
      }
      const double V = dij[3] / 3.0 + dji[3] / 3.0;
      double out[14];
#pragma unroll
      for (int r = 0; r < 14; r++) out[r] = 0.0;
      if (ninside_pair > 0 && V > 0) {
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
      if (lane == 0) {
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
      __syncwarp();
    }
  }
  if (lane == 0) {
    atomicAdd(&A.counters[0], n_pairs);
    atomicAdd(&A.counters[1], n_trans);
    atomicAdd(&A.counters[2], n_eval);
    atomicAdd(&A.counters[3], n_inside);
    atomicAdd(&A.counters[4], n_gh);
  }
}

inline size_t pair_warp_smem_bytes(int total_terms, int nw, bool smem_tables) {
  size_t b = 0;
  if (smem_tables) b += (size_t)total_terms * (sizeof(double2) + sizeof(double));
  b += (size_t)nw * 26 * sizeof(double) + (size_t)nw * 128 * sizeof(unsigned short);
  return (b + 15) & ~(size_t)15;
}

}  // namespace shgpu
