// pair_split_kernels.cuh — three-kernel form of the SPHERHARM pair phase (default), sm_100a.
//
// Same arithmetic contract and the same node decisions as pair_warp_kernel.cuh / the oracle; the work is
// re-cut so that each kernel runs at the occupancy its bottleneck wants:
//   A  pair_cull_kernel   (latency bound: node-table gathers from L2)  one warp per pair, <= 64 registers,
//      up to 64 warps/SM.  Window -> exact transform -> bounding-sphere test -> direction-cell bound ->
//      survivor RECORDS (s-vector in b's frame + node index, 32 B) compacted in shared memory and appended
//      as one contiguous run per (pair, direction) to the pool of the TARGET shape (one atomicAdd per run).
//      Pairs with more than SPLIT_CAP survivors in a direction ("deep contacts") go to a list that the
//      fused warp kernel processes afterwards.
//   B  pair_eval_kernel   (FP64 FMA-pipe bound)  every pool holds records against ONE shape, so warps are
//      full and shape-uniform whatever pair the records came from: 64 records per warp-task, two points
//      per lane sharing the coefficient loads, the shape's folded table staged once per CTA.  Writes the
//      inside flag into the record.
//   C  pair_reduce_kernel (small)  one warp per pair walks its two record runs in order, accumulates the
//      inside nodes (fixed order -> bitwise reproducible), applies the contact law, writes the outputs.
// No floating-point atomics anywhere; pool offsets vary from run to run but never enter the arithmetic.
#pragma once
#include "pair_kernel.cuh"

namespace shgpu {

#define SPLIT_CAP 88

struct __align__(16) SurvRec {
  double s0, s1, s2;
  int k;       // node index in a's table
  int flag;    // 0 = outside, 1 = inside (set by B), 2 = inside b's inscribed sphere (set by A)
};

struct SplitArgs {
  SurvRec *pool;                  // all shapes' pools in one buffer
  unsigned char *pool_flag;       // inside flag per record (own array: a flag store must not dirty the 32 B record)
  const long long *pool_base;     // [nshape] first record of the shape's pool
  const long long *pool_cap;      // [nshape]
  unsigned long long *pool_count; // [nshape] records appended this step (may exceed cap -> host grows and reruns)
  long long *pd_off;              // [2P] first record of the run (absolute index into pool)
  int *pd_cnt;                    // [2P] run length; -1 = pair is on the deep-contact list
  int *big_list;                  // [P]
  int *nbig;
  int *overflow;
};

// ---- conservative window on a's node grid (FP32 with margins), shared by the cull and cache-build kernels.
// Rb = radius of the sphere around b's origin that a node must enter (rmax_b, or rmax_b + delta for the cache).
struct NodeWindow { bool skip; float cosA, xe, se, phie; };

__device__ __forceinline__ NodeWindow make_window(const DevShape &sa, double Rb, const double *x0) {
  NodeWindow w;
  const double e0 = 2.0 * x0[0], e1 = 2.0 * x0[1], e2 = 2.0 * x0[2];   // b's origin in a's frame
  const double D2 = e0 * e0 + e1 * e1 + e2 * e2, D = sqrt(D2);
  w.skip = D >= (sa.rmax + Rb) * (1.0 + 1e-9);
  w.cosA = -2.0f; w.xe = 1.0f; w.se = 0.0f; w.phie = 0.0f;
  if (!w.skip && D > 1e-9) {
    const double q = D2 - Rb * Rb;
    double rc = sa.rmin;
    if (q > 0) rc = fmin(fmax(sqrt(q), sa.rmin), sa.rmax);
    const double g = (rc * rc + q) / (2.0 * rc * D);
    w.cosA = (float)g - 3e-5f;
    if (w.cosA >= 1.0f) w.skip = true;
    w.xe = fminf(1.0f, fmaxf(-1.0f, (float)(e2 / D)));
    w.se = sqrtf(fmaxf(0.0f, 1.0f - w.xe * w.xe));
    w.phie = atan2f((float)e1, (float)e0);
    if (w.phie < 0.0f) w.phie += 6.2831853f;
  }
  return w;
}

__device__ __forceinline__ void classify_row(const DevShape &sa, const NodeWindow &w, int row, int &c0, int &ccount) {
  c0 = 0; ccount = 0;
  const int nph = sa.n_phi;
  if (row >= sa.n_theta) return;
  if (w.cosA <= -1.0f) { ccount = nph; return; }
  const float xa = sa.row_x[row];
  const float sarow = sqrtf(fmaxf(0.0f, 1.0f - xa * xa));
  const float ss = sarow * w.se, xx = xa * w.xe;
  if (xx + ss < w.cosA) return;
  const float cd = (ss > 1e-12f) ? (w.cosA - xx) / ss : -2.0f;
  if (cd <= -1.0f) { ccount = nph; return; }
  const float inv_dphi = (float)nph * 0.15915494f;
  const float dl = acosf(fminf(cd, 1.0f)) + 2e-4f;
  const int b0 = (int)ceilf((w.phie - dl) * inv_dphi - 0.5f);
  const int b1 = (int)floorf((w.phie + dl) * inv_dphi - 0.5f);
  ccount = b1 - b0 + 1;
  if (ccount >= nph) { ccount = nph; c0 = 0; }
  else if (ccount > 0) { c0 = b0 % nph; if (c0 < 0) c0 += nph; }
  else ccount = 0;
}

// cube-map direction cell of a vector (FP32); the tables are conservative across cell borders
__device__ __forceinline__ int cube_cell(float f0, float f1, float f2, int cn) {
  const float ax = fabsf(f0), ay = fabsf(f1), az = fabsf(f2);
  int face; float ma, uu, vv;
  if (ax >= ay && ax >= az) { face = f0 > 0 ? 0 : 1; ma = ax; uu = f1; vv = f2; }
  else if (ay >= az) { face = f1 > 0 ? 2 : 3; ma = ay; uu = f0; vv = f2; }
  else { face = f2 > 0 ? 4 : 5; ma = az; uu = f0; vv = f1; }
  const float im = 1.0f / fmaxf(ma, 1e-30f), hn = 0.5f * (float)cn;
  const int iu = min(cn - 1, max(0, (int)((uu * im + 1.0f) * hn)));
  const int iv = min(cn - 1, max(0, (int)((vv * im + 1.0f) * hn)));
  return (face * cn + iu) * cn + iv;
}

// relative pose of direction a -> b into the warp's shared scratch: M[9], t[3], x0[3] (same fma chains as the oracle)
__device__ __forceinline__ void pose_to_smem(const PairArgs &A, int a, int b, double dd0, double dd1, double dd2, int lane,
                                             double *pose) {
  const int st = A.stride;
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
}

#define CACHE_CAP 256

struct CacheArgs {
  unsigned short *pool;            // candidate node indices of all (pair, direction) runs
  long long *off;                  // [2P]
  int *cnt;                        // [2P]  -1 = not cached (too many candidates): the cull falls back to the window
  unsigned long long *count;       // appended entries (may exceed cap -> host grows and rebuilds)
  long long cap;
  int *overflow;
  const int *invalid;              // device flag raised by cache_check_kernel: the cull ignores the cache when set
  int enabled;
};

// ---- candidate cache build (on neighbor rebuilds and when a particle has used up its displacement margin):
// window + FP32 tests against the INFLATED bounds (rmax_b + delta; cube_w2), no FP64 work.
template <int WPB>
__global__ void __launch_bounds__(WPB * 32, 4) pair_cache_build_kernel(PairArgs A, CacheArgs C) {
  __shared__ unsigned short s_buf[WPB][CACHE_CAP];
  __shared__ double s_pose[WPB][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * WPB + warp;
  if (p >= A.npairs) return;
  const int st = A.stride;
  const int i = A.pair_i[p], j = A.pair_j[p];
  double d[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double dk = A.c[k * st + i] - A.c[k * st + j];
    if (A.periodic[k]) dk = dk - A.boxlen[k] * rint(dk / A.boxlen[k]);
    d[k] = dk;
  }
  const int shp_i = A.shape[i], shp_j = A.shape[j];
  double *pose = s_pose[warp];
  unsigned short *buf = s_buf[warp];
  for (int dir = 0; dir < 2; dir++) {
    const int a = dir ? j : i, b = dir ? i : j;
    const DevShape &sa = A.shapes[dir ? shp_j : shp_i];
    const DevShape &sb = A.shapes[dir ? shp_i : shp_j];
    const double sgn = dir ? -1.0 : 1.0;
    pose_to_smem(A, a, b, sgn * d[0], sgn * d[1], sgn * d[2], lane, pose);
    const double *M = pose, *t = pose + 9, *x0 = pose + 12;
    const double delta = sb.cache_delta, Rb = sb.rmax + delta;
    const NodeWindow w = make_window(sa, Rb, x0);
    const float *__restrict__ pf = sa.pf;
    const float *__restrict__ cube = sb.cube_w2;
    const int cn = sb.cube_n, nq = sa.nq, nph = sa.n_phi, nth = sa.n_theta;
    const float fM0 = (float)M[0], fM1 = (float)M[1], fM2 = (float)M[2], fM3 = (float)M[3], fM4 = (float)M[4],
                fM5 = (float)M[5], fM6 = (float)M[6], fM7 = (float)M[7], fM8 = (float)M[8];
    const float ft0 = (float)t[0], ft1 = (float)t[1], ft2 = (float)t[2];
    const float fR2 = (float)(Rb * Rb) + 1e-4f;
    const float fin = (float)((sb.rmin + 2.0 * delta) * (sb.rmin + 2.0 * delta)) + 1e-4f;
    int n = 0;
    bool over = false;
    if (!w.skip) {
      for (int rb = 0; rb < nth && !over; rb += 32) {
        int c0, ccount;
        classify_row(sa, w, rb + lane, c0, ccount);
        unsigned rows = __ballot_sync(0xffffffffu, ccount > 0);
        while (rows && !over) {
          const int rl = __ffs(rows) - 1;
          rows &= rows - 1;
          const int rc0 = __shfl_sync(0xffffffffu, c0, rl), rcount = __shfl_sync(0xffffffffu, ccount, rl);
          const int rowbase = (rb + rl) * nph;
          for (int cb = 0; cb < rcount; cb += 32) {
            const int cc = cb + lane;
            bool pass = false;
            int k = 0;
            if (cc < rcount) {
              int col = rc0 + cc;
              if (col >= nph) col -= nph;
              k = rowbase + col;
              const float q0 = pf[k], q1 = pf[nq + k], q2 = pf[2 * nq + k];
              const float f0 = fmaf(fM2, q2, fmaf(fM1, q1, fmaf(fM0, q0, ft0)));
              const float f1 = fmaf(fM5, q2, fmaf(fM4, q1, fmaf(fM3, q0, ft1)));
              const float f2 = fmaf(fM8, q2, fmaf(fM7, q1, fmaf(fM6, q0, ft2)));
              const float r2 = fmaf(f2, f2, fmaf(f1, f1, f0 * f0));
              if (r2 < fR2) pass = (r2 < fin) || (r2 < __ldg(&cube[cube_cell(f0, f1, f2, cn)]) + 1e-4f);
            }
            const unsigned pm = __ballot_sync(0xffffffffu, pass);
            if (pm) {
              const int nnew = __popc(pm);
              if (n + nnew > CACHE_CAP) { over = true; break; }
              if (pass) buf[n + __popc(pm & ((1u << lane) - 1u))] = (unsigned short)k;
              n += nnew;
            }
          }
        }
      }
    }
    __syncwarp();
    long long off = 0;
    if (over) n = -1;
    if (lane == 0 && n > 0) off = (long long)atomicAdd(C.count, (unsigned long long)n);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (n > 0) {
      if (off + n <= C.cap) { for (int r = lane; r < n; r += 32) C.pool[off + r] = buf[r]; }
      else { if (lane == 0) *C.overflow = 1; n = -1; }
    }
    if (lane == 0) { C.off[2 * p + dir] = off; C.cnt[2 * p + dir] = n; }
    __syncwarp();
  }
}

// ---- A: cull.  Per (pair, direction): candidates (from the cache, else from the window + FP32 pre-cull) ->
// exact FP64 stage on full warps -> survivor records -> one contiguous run in the target shape's pool.
template <int WPB>
__global__ void __launch_bounds__(WPB * 32, 32 / WPB) pair_cull_kernel(PairArgs A, SplitArgs S, CacheArgs C, int use_bounds) {
  __shared__ __align__(16) SurvRec s_rec[WPB][2][SPLIT_CAP];
  __shared__ double s_pose[WPB][16];
  __shared__ unsigned short s_cand[WPB][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * WPB + warp;
  if (p < A.npairs) {
  const int st = A.stride;
  const int i = A.pair_i[p], j = A.pair_j[p];
  double d[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double dk = A.c[k * st + i] - A.c[k * st + j];
    if (A.periodic[k]) dk = dk - A.boxlen[k] * rint(dk / A.boxlen[k]);
    d[k] = dk;
  }
  const int shp_i = A.shape[i], shp_j = A.shape[j];
  double *pose = s_pose[warp];
  int cnt[2] = {0, 0};
  bool big = false;
  unsigned long long n_trans = 0;
  bool cached = false;
  int ccnt[2] = {0, 0};
  if (C.enabled && *C.invalid == 0) {
    ccnt[0] = C.cnt[2 * p]; ccnt[1] = C.cnt[2 * p + 1];
    cached = ccnt[0] >= 0 && ccnt[1] >= 0;
  }

  for (int dir = 0; dir < 2 && !big; dir++) {
    const int a = dir ? j : i, b = dir ? i : j;
    const DevShape &sa = A.shapes[dir ? shp_j : shp_i];
    const DevShape &sb = A.shapes[dir ? shp_i : shp_j];
    const double sgn = dir ? -1.0 : 1.0;
    if (cached && ccnt[dir] == 0) { cnt[dir] = 0; continue; }
    pose_to_smem(A, a, b, sgn * d[0], sgn * d[1], sgn * d[2], lane, pose);
    const double *M = pose, *t = pose + 9, *x0 = pose + 12;
    const double rmax2 = sb.rmax2, rmin2 = sb.rmin2;
    const double *__restrict__ nodes = sa.px;
    const int nq = sa.nq;
    const float *__restrict__ cube = sb.cube_b2;
    const int cn = sb.cube_n;
    SurvRec *buf = s_rec[warp][dir];
    unsigned short *cand = s_cand[warp];
    int queued = 0;

    // exact FP64 stage: transform, bounding sphere, inscribed sphere, direction-cell bound
    auto exact_stage = [&](bool valid, int k) {
      int flag = -1;
      double s0 = 0, s1 = 0, s2 = 0;
      if (valid) {
        const double p0 = nodes[k], p1 = nodes[nq + k], p2 = nodes[2 * nq + k];
        s0 = fma(M[0], p0, t[0]); s0 = fma(M[1], p1, s0); s0 = fma(M[2], p2, s0);
        s1 = fma(M[3], p0, t[1]); s1 = fma(M[4], p1, s1); s1 = fma(M[5], p2, s1);
        s2 = fma(M[6], p0, t[2]); s2 = fma(M[7], p1, s2); s2 = fma(M[8], p2, s2);
        const double rho2 = fma(s2, s2, fma(s1, s1, s0 * s0));
        if (rho2 < rmax2) {
          if (rho2 <= rmin2) flag = 2;
          else if (use_bounds) { if (rho2 < (double)__ldg(&cube[cube_cell((float)s0, (float)s1, (float)s2, cn)])) flag = 0; }
          else flag = 0;
        }
      }
      const unsigned sm = __ballot_sync(0xffffffffu, flag >= 0);
      if (sm) {
        const int nnew = __popc(sm);
        if (queued + nnew > SPLIT_CAP) { big = true; return; }
        if (flag >= 0) {
          SurvRec r; r.s0 = s0; r.s1 = s1; r.s2 = s2; r.k = k; r.flag = flag;
          buf[queued + __popc(sm & ((1u << lane) - 1u))] = r;
        }
        queued += nnew;
      }
    };

    if (cached) {
      const unsigned short *__restrict__ cl = C.pool + C.off[2 * p + dir];
      const int m = ccnt[dir];
      for (int g = 0; g < m && !big; g += 32) {
        const bool valid = g + lane < m;
        exact_stage(valid, valid ? (int)cl[g + lane] : 0);
      }
      n_trans += m;
    } else {
      const NodeWindow w = make_window(sa, sb.rmax, x0);
      if (!w.skip) {
        const float *__restrict__ pf = sa.pf;
        const float fM0 = (float)M[0], fM1 = (float)M[1], fM2 = (float)M[2], fM3 = (float)M[3], fM4 = (float)M[4],
                    fM5 = (float)M[5], fM6 = (float)M[6], fM7 = (float)M[7], fM8 = (float)M[8];
        const float ft0 = (float)t[0], ft1 = (float)t[1], ft2 = (float)t[2];
        const float frmax2 = (float)rmax2 + 1e-4f;
        const int nth = sa.n_theta, nph = sa.n_phi;
        int ncand = 0;
        for (int rb = 0; rb < nth && !big; rb += 32) {
          int c0, ccount;
          classify_row(sa, w, rb + lane, c0, ccount);
          unsigned rows = __ballot_sync(0xffffffffu, ccount > 0);
          while (rows && !big) {
            const int rl = __ffs(rows) - 1;
            rows &= rows - 1;
            const int rc0 = __shfl_sync(0xffffffffu, c0, rl), rcount = __shfl_sync(0xffffffffu, ccount, rl);
            const int rowbase = (rb + rl) * nph;
            for (int cb = 0; cb < rcount && !big; cb += 32) {
              const int cc = cb + lane;
              bool pass = false;
              int k = 0;
              if (cc < rcount) {
                int col = rc0 + cc;
                if (col >= nph) col -= nph;
                k = rowbase + col;
                const float q0 = pf[k], q1 = pf[nq + k], q2 = pf[2 * nq + k];
                const float f0 = fmaf(fM2, q2, fmaf(fM1, q1, fmaf(fM0, q0, ft0)));
                const float f1 = fmaf(fM5, q2, fmaf(fM4, q1, fmaf(fM3, q0, ft1)));
                const float f2 = fmaf(fM8, q2, fmaf(fM7, q1, fmaf(fM6, q0, ft2)));
                const float r2 = fmaf(f2, f2, fmaf(f1, f1, f0 * f0));
                if (r2 < frmax2) pass = !use_bounds || r2 < __ldg(&cube[cube_cell(f0, f1, f2, cn)]) + 1e-4f;
              }
              n_trans += min(32, rcount - cb);
              const unsigned pm = __ballot_sync(0xffffffffu, pass);
              if (pm) {
                if (pass) cand[ncand + __popc(pm & ((1u << lane) - 1u))] = (unsigned short)k;
                ncand += __popc(pm);
                __syncwarp();
                if (ncand >= 32) {
                  exact_stage(true, (int)cand[lane]);
                  __syncwarp();
                  const int rem = ncand - 32;
                  unsigned short mv = 0;
                  if (lane < rem) mv = cand[32 + lane];
                  __syncwarp();
                  if (lane < rem) cand[lane] = mv;
                  ncand = rem;
                  __syncwarp();
                }
              }
            }
          }
        }
        if (!big && ncand > 0) { exact_stage(lane < ncand, lane < ncand ? (int)cand[lane] : 0); __syncwarp(); }
      }
    }
    cnt[dir] = queued;
  }
  __syncwarp();
  if (big) {
    if (lane == 0) {
      S.pd_cnt[2 * p] = -1; S.pd_cnt[2 * p + 1] = -1;
      S.big_list[atomicAdd(S.nbig, 1)] = p;
    }
    // counters of a deep pair are accumulated by the fused kernel that re-does it
  } else {
#pragma unroll
  for (int dir = 0; dir < 2; dir++) {
    const int sb_id = dir ? shp_i : shp_j;
    long long off = 0;
    if (lane == 0 && cnt[dir] > 0) off = (long long)atomicAdd(&S.pool_count[sb_id], (unsigned long long)cnt[dir]);
    off = __shfl_sync(0xffffffffu, off, 0);
    const bool fits = off + cnt[dir] <= S.pool_cap[sb_id];
    const long long base = S.pool_base[sb_id] + off;
    if (fits) {
      for (int r = lane; r < cnt[dir]; r += 32) {
        S.pool[base + r] = s_rec[warp][dir][r];
        S.pool_flag[base + r] = (unsigned char)s_rec[warp][dir][r].flag;
      }
    } else if (lane == 0) *S.overflow = 1;
    if (lane == 0) { S.pd_off[2 * p + dir] = base; S.pd_cnt[2 * p + dir] = fits ? cnt[dir] : 0; }
  }
  if (lane == 0) {   // no block barrier here: warps must retire independently (pairs differ a lot in length)
    atomicAdd(&A.counters[0], 1ull);
    atomicAdd(&A.counters[1], n_trans);
    if (j >= A.nlocal) atomicAdd(&A.counters[4], 1ull);
  }
  }  // !big
  }  // p < npairs
}

// ---- B: evaluate pooled records.  blockIdx -> (shape, first record) through blk_start (prefix of CTA counts)
struct EvalPlan {
  int nshape;
  int blk_start[SH_MAX_SHAPES + 1];
  long long count[SH_MAX_SHAPES];
};

template <int WPB>
__global__ void __launch_bounds__(WPB * 32) pair_eval_kernel(const DevShape *shapes, SplitArgs S, EvalPlan P,
                                                             unsigned long long *counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int s = 0;
  while (s + 1 < P.nshape && (int)blockIdx.x >= P.blk_start[s + 1]) s++;
  const DevShape &sh = shapes[s];
  double2 *s_ab = reinterpret_cast<double2 *>(smem_raw);
  double *s_Ap = reinterpret_cast<double *>(s_ab + sh.nterms4 + 4);
  for (int t = threadIdx.x; t < sh.nterms4 + 4; t += WPB * 32) { s_ab[t] = sh.ab[t]; s_Ap[t] = sh.Ap[t]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ unsigned long long s_nev;
  if (threadIdx.x == 0) s_nev = 0;
  __syncthreads();
  const long long n = P.count[s];
  const long long r0 = ((long long)(blockIdx.x - P.blk_start[s]) * WPB + warp) * 64;
  if (r0 < n) {
  SurvRec *rec = S.pool + S.pool_base[s];
  unsigned char *flg = S.pool_flag + S.pool_base[s];
  const long long ia = r0 + lane, ib = r0 + 32 + lane;
  const bool va = ia < n, vb = ib < n;
  const int L = sh.lmax;
  int nev = 0;
  if (__ballot_sync(0xffffffffu, vb)) {
    const long long ja = va ? ia : r0, jb = vb ? ib : (va ? ia : r0);
    const SurvRec ra = rec[ja], rb = rec[jb];
    const double sA[3] = {ra.s0, ra.s1, ra.s2}, sB[3] = {rb.s0, rb.s1, rb.s2};
    const double rhoA2 = fma(sA[2], sA[2], fma(sA[1], sA[1], sA[0] * sA[0]));
    const double rhoB2 = fma(sB[2], sB[2], fma(sB[1], sB[1], sB[0] * sB[0]));
    double rhoA, rhoB, rA, rB;
    sh_radius_folded_x2(L, s_Ap, s_ab, sA, rhoA2, sB, rhoB2, rhoA, rhoB, rA, rB);
    if (va && ra.flag != 2) { flg[ia] = rhoA < rA ? 1 : 0; nev++; }
    if (vb && rb.flag != 2) { flg[ib] = rhoB < rB ? 1 : 0; nev++; }
  } else if (va) {
    const SurvRec ra = rec[ia];
    const double rho2 = fma(ra.s2, ra.s2, fma(ra.s1, ra.s1, ra.s0 * ra.s0));
    double rho;
    const double r = sh_radius_folded(L, s_Ap, s_ab, ra.s0, ra.s1, ra.s2, rho2, rho);
    if (ra.flag != 2) { flg[ia] = rho < r ? 1 : 0; nev++; }
  }
  nev = __reduce_add_sync(0xffffffffu, nev);   // records preset by A (flag 2) are not "evaluated"
  if (lane == 0 && nev) atomicAdd(&s_nev, (unsigned long long)nev);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_nev) { atomicAdd(&counters[2], s_nev); atomicAdd(&counters[5], s_nev); }
}

// ---- C: per-pair reduction + contact law.  One THREAD per pair: the two record runs are
// short (tens of records, a few of them inside) and are summed sequentially in record order.
__global__ void __launch_bounds__(128, 6) pair_reduce_kernel(PairArgs A, SplitArgs S) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.npairs) return;
  if (S.pd_cnt[2 * p] < 0) return;   // deep contact: done by the fused kernel
  const int st = A.stride;
  const int i = A.pair_i[p], j = A.pair_j[p];
  double d[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double dk = A.c[k * st + i] - A.c[k * st + j];
    if (A.periodic[k]) dk = dk - A.boxlen[k] * rint(dk / A.boxlen[k]);
    d[k] = dk;
  }
  const int shp_i = A.shape[i], shp_j = A.shape[j];
  double out[14];
#pragma unroll
  for (int r = 0; r < 14; r++) out[r] = 0.0;
  double Ss[2][3], Ts[2][3], Gs[2][3], Asum[2];
  int ninside_pair = 0;
#pragma unroll
  for (int dir = 0; dir < 2; dir++) {
    const int a = dir ? j : i;
    const DevShape &sa = A.shapes[dir ? shp_j : shp_i];
    const double sgn = dir ? -1.0 : 1.0;
    const long long off = S.pd_off[2 * p + dir];
    const int cnt = S.pd_cnt[2 * p + dir];
    double S0 = 0, S1 = 0, S2 = 0, Av = 0, T0 = 0, T1 = 0, T2 = 0, G0 = 0, G1 = 0, G2 = 0;
    double Ra[9];
#pragma unroll
    for (int e = 0; e < 9; e++) Ra[e] = A.Rs[e * st + a];
    if (cnt > 0) {
      double x0[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double hx = Ra[0 + r] * (sgn * d[0]);
        hx = fma(Ra[3 + r], sgn * d[1], hx);
        hx = fma(Ra[6 + r], sgn * d[2], hx);
        x0[r] = -0.5 * hx;
      }
      const double *__restrict__ nodes = sa.px;
      const int nq = sa.nq;
      for (int r = 0; r < cnt; r++) {
        if (S.pool_flag[off + r]) {
          const int k = S.pool[off + r].k;
          const double p0 = nodes[k], p1 = nodes[nq + k], p2 = nodes[2 * nq + k];
          const double n0 = nodes[3 * nq + k], n1 = nodes[4 * nq + k], n2 = nodes[5 * nq + k];
          const double dp0 = p0 - x0[0], dp1 = p1 - x0[1], dp2 = p2 - x0[2];
          const double dn = fma(dp2, n2, fma(dp1, n1, dp0 * n0));
          S0 += n0; S1 += n1; S2 += n2;
          Av += dn;
          T0 += fma(p1, n2, -(p2 * n1));
          T1 += fma(p2, n0, -(p0 * n2));
          T2 += fma(p0, n1, -(p1 * n0));
          G0 = fma(dp0, dn, G0); G1 = fma(dp1, dn, G1); G2 = fma(dp2, dn, G2);
          ninside_pair++;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
      Ss[dir][r] = Ra[3 * r] * S0 + Ra[3 * r + 1] * S1 + Ra[3 * r + 2] * S2;
      Ts[dir][r] = Ra[3 * r] * T0 + Ra[3 * r + 1] * T1 + Ra[3 * r + 2] * T2;
      Gs[dir][r] = 0.25 * (Ra[3 * r] * G0 + Ra[3 * r + 1] * G1 + Ra[3 * r + 2] * G2);
    }
    Asum[dir] = Av;
  }
  const double V = Asum[0] / 3.0 + Asum[1] / 3.0;
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
    const double *Sij = Ss[0], *Sji = Ss[1], *Tij = Ts[0], *Tji = Ts[1];
    const double Ti[3] = {Tij[0] + (li[1] * Sij[2] - li[2] * Sij[1]), Tij[1] + (li[2] * Sij[0] - li[0] * Sij[2]),
                          Tij[2] + (li[0] * Sij[1] - li[1] * Sij[0])};
    const double Tj[3] = {Tji[0] + (lj[1] * Sji[2] - lj[2] * Sji[1]), Tji[1] + (lj[2] * Sji[0] - lj[0] * Sji[2]),
                          Tji[2] + (lj[0] * Sji[1] - lj[1] * Sji[0])};
#pragma unroll
    for (int r = 0; r < 3; r++) {
      out[2 + r] = -pr * (0.5 * (Sij[r] - Sji[r]));
      out[5 + r] = -pr * Ti[r];
      out[8 + r] = -pr * Tj[r];
      out[11 + r] = (A.c[r * st + i] - 0.5 * d[r]) + (Gs[0][r] + Gs[1][r]) / V;
    }
  }
#pragma unroll
  for (int r = 0; r < 14; r++) A.pres[(size_t)r * A.pres_stride + p] = out[r];   // coalesced across pairs
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
  // inside-node counter: one atomic per warp
  const int tot = __reduce_add_sync(__activemask(), ninside_pair);
  if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1) && tot) atomicAdd(&A.counters[3], (unsigned long long)tot);
}

}  // namespace shgpu
