// dfma_probe.cu — microbenchmarks that size the SH-evaluation inner loop on B200 (sm_100a).
//   1. DFMA dependent-issue latency and throughput vs ILP / warps per SM
//   2. the folded-recurrence inner loop in several structures (points per thread, prefetch,
//      block-boundary handling) at several warp counts, tables in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o build/dfma_probe tools/dfma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../lammps-spherharm_b200/csrc/device_math.cuh"

using namespace shgpu;

// (kept here only: measured slower than the plain nested loop)
// Software-pipelined form of sh_radius_folded for tables padded to a multiple of 4 records plus 4
// (zero records): the coefficient records of the NEXT four terms are fetched while the current
// four are consumed, so the shared-memory latency is off the dependent chain.  Every term runs the
// same body; a block (fixed m) starts from the state q1 = 0, q2 = -1, C = S = 0, which makes the
// generic body reproduce the special first two terms exactly:
//   l = m   : tx = 0*x,  q = fma(tx, 0, 1) = 1,        C = fma(a, 1, 0) = a
//   l = m+1 : q = fma(Ap*x, 1, -0) = Ap*x (the rounded product), as in sh_radius_folded.
// Bit-identical to sh_radius_folded (and to the oracle) up to the sign of a zero accumulator.
__device__ __forceinline__ double sh_radius_folded_pipe(int L, int nterms4, const double *__restrict__ Ap,
                                                        const double2 *__restrict__ ab, double s0, double s1,
                                                        double s2, double rho2, double &rho_out) {
  const double rho = sqrt(rho2);
  const double inv = 1.0 / rho;
  const double x = s2 * inv, zx = s0 * inv, zy = s1 * inv;
  double u = 1.0, v = 0.0, r = 0.0, C = 0.0, S = 0.0, q1 = 0.0, q2 = -1.0;
  int m = 0, bend = L;
  double nA0 = Ap[0], nA1 = Ap[1], nA2 = Ap[2], nA3 = Ap[3];
  double2 nC0 = ab[0], nC1 = ab[1], nC2 = ab[2], nC3 = ab[3];
#pragma unroll 1
  for (int idx = 0; idx < nterms4; idx += 4) {
    const double cA0 = nA0, cA1 = nA1, cA2 = nA2, cA3 = nA3;
    const double2 cC0 = nC0, cC1 = nC1, cC2 = nC2, cC3 = nC3;
    nA0 = Ap[idx + 4]; nA1 = Ap[idx + 5]; nA2 = Ap[idx + 6]; nA3 = Ap[idx + 7];
    nC0 = ab[idx + 4]; nC1 = ab[idx + 5]; nC2 = ab[idx + 6]; nC3 = ab[idx + 7];
#define SH_TERM(cA, cC, J)                                                      \
    {                                                                           \
      const double tx = (cA) * x;                                               \
      const double q = fma(tx, q1, -q2);                                        \
      C = fma((cC).x, q, C); S = fma((cC).y, q, S);                             \
      q2 = q1; q1 = q;                                                          \
      if (idx + (J) == bend) {                                                  \
        r = fma(u, C, r); r = fma(v, S, r);                                     \
        m++; bend += L + 1 - m;                                                 \
        const double t1 = v * zy, un = fma(u, zx, -t1);                         \
        const double t2 = v * zx, vn = fma(u, zy, t2);                          \
        u = un; v = vn; C = 0.0; S = 0.0; q1 = 0.0; q2 = -1.0;                  \
      }                                                                         \
    }
    SH_TERM(cA0, cC0, 0) SH_TERM(cA1, cC1, 1) SH_TERM(cA2, cC2, 2) SH_TERM(cA3, cC3, 3)
#undef SH_TERM
  }
  rho_out = rho;
  return r;
}



#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void dfma_chain(double *out, int iters, double a, double b) {
  double r[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) r[k] = threadIdx.x * 1e-3 + k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int k = 0; k < ILP; k++) r[k] = fma(r[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += r[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- inner-loop variants. Table: Ap[T], ab[T] in smem (padded). Each thread evaluates npts points.
// variant 0: reference nested loops (sh_radius_folded), 1 point
// variant 1: flat pipelined (sh_radius_folded_pipe), 1 point
// variant 2: nested loops, 2 points per thread sharing the coefficient loads
// variant 3: nested loops, 1 point, explicit prefetch of the next 4 terms within a block
// variant 4: like 2 with explicit prefetch
template <int VAR>
__global__ void eval_loop(double *out, int L, int T4, const double *gAp, const double2 *gab, int npts) {
  extern __shared__ __align__(16) unsigned char sm[];
  double2 *s_ab = reinterpret_cast<double2 *>(sm);
  double *s_Ap = reinterpret_cast<double *>(s_ab + T4 + 8);
  for (int t = threadIdx.x; t < T4 + 8; t += blockDim.x) { s_ab[t] = gab[t]; s_Ap[t] = gAp[t]; }
  __syncthreads();
  double acc = 0;
  const double sx = 0.3 + 1e-3 * threadIdx.x, sy = 0.5, sz = 0.2 + 1e-4 * blockIdx.x;
  for (int p = 0; p < npts; p++) {
    const double s0 = sx + 1e-3 * p, s1 = sy, s2 = sz;
    const double rho2 = fma(s2, s2, fma(s1, s1, s0 * s0));
    double rho;
    if (VAR == 0) acc += sh_radius_folded(L, s_Ap, s_ab, s0, s1, s2, rho2, rho);
    else if (VAR == 1) acc += sh_radius_folded_pipe(L, T4, s_Ap, s_ab, s0, s1, s2, rho2, rho);
    else if (VAR == 2 || VAR == 4) {
      // two points per thread
      const double t0 = s0 + 0.37, t1 = s1 - 0.11, t2 = s2 + 0.05;
      const double rhoB2 = fma(t2, t2, fma(t1, t1, t0 * t0));
      const double rA = sqrt(rho2), iA = 1.0 / rA, rB = sqrt(rhoB2), iB = 1.0 / rB;
      const double xA = s2 * iA, zxA = s0 * iA, zyA = s1 * iA, xB = t2 * iB, zxB = t0 * iB, zyB = t1 * iB;
      double uA = 1, vA = 0, rrA = 0, uB = 1, vB = 0, rrB = 0;
      int base = 0;
      for (int m = 0; m <= L; m++) {
        if (m > 0) {
          double a1 = vA * zyA, un = fma(uA, zxA, -a1), a2 = vA * zxA, vn = fma(uA, zyA, a2); uA = un; vA = vn;
          double b1 = vB * zyB, wn = fma(uB, zxB, -b1), b2 = vB * zxB, yn = fma(uB, zyB, b2); uB = wn; vB = yn;
        }
        const double2 c0 = s_ab[base];
        double CA = c0.x, SA = c0.y, CB = c0.x, SB = c0.y;
        const int len = L - m;
        if (len >= 1) {
          const double ap1 = s_Ap[base + 1];
          double qA1 = ap1 * xA, qA2 = 1.0, qB1 = ap1 * xB, qB2 = 1.0;
          const double2 c1 = s_ab[base + 1];
          CA = fma(c1.x, qA1, CA); SA = fma(c1.y, qA1, SA); CB = fma(c1.x, qB1, CB); SB = fma(c1.y, qB1, SB);
          if (VAR == 2) {
#pragma unroll 4
            for (int i = 2; i <= len; i++) {
              const double ap = s_Ap[base + i];
              const double2 ci = s_ab[base + i];
              const double txA = ap * xA, txB = ap * xB;
              const double qA = fma(txA, qA1, -qA2), qB = fma(txB, qB1, -qB2);
              CA = fma(ci.x, qA, CA); SA = fma(ci.y, qA, SA); CB = fma(ci.x, qB, CB); SB = fma(ci.y, qB, SB);
              qA2 = qA1; qA1 = qA; qB2 = qB1; qB1 = qB;
            }
          } else {
            int i = 2;
            double nap0 = s_Ap[base + 2], nap1 = s_Ap[base + 3];
            double2 nc0 = s_ab[base + 2], nc1 = s_ab[base + 3];
            for (; i + 1 <= len; i += 2) {
              const double ap0 = nap0, ap1b = nap1; const double2 ci0 = nc0, ci1 = nc1;
              nap0 = s_Ap[base + i + 2]; nap1 = s_Ap[base + i + 3]; nc0 = s_ab[base + i + 2]; nc1 = s_ab[base + i + 3];
              {
                const double txA = ap0 * xA, txB = ap0 * xB;
                const double qA = fma(txA, qA1, -qA2), qB = fma(txB, qB1, -qB2);
                CA = fma(ci0.x, qA, CA); SA = fma(ci0.y, qA, SA); CB = fma(ci0.x, qB, CB); SB = fma(ci0.y, qB, SB);
                qA2 = qA1; qA1 = qA; qB2 = qB1; qB1 = qB;
              }
              {
                const double txA = ap1b * xA, txB = ap1b * xB;
                const double qA = fma(txA, qA1, -qA2), qB = fma(txB, qB1, -qB2);
                CA = fma(ci1.x, qA, CA); SA = fma(ci1.y, qA, SA); CB = fma(ci1.x, qB, CB); SB = fma(ci1.y, qB, SB);
                qA2 = qA1; qA1 = qA; qB2 = qB1; qB1 = qB;
              }
            }
            if (i <= len) {
              const double txA = nap0 * xA, txB = nap0 * xB;
              const double qA = fma(txA, qA1, -qA2), qB = fma(txB, qB1, -qB2);
              CA = fma(nc0.x, qA, CA); SA = fma(nc0.y, qA, SA); CB = fma(nc0.x, qB, CB); SB = fma(nc0.y, qB, SB);
            }
          }
        }
        rrA = fma(uA, CA, rrA); rrA = fma(vA, SA, rrA); rrB = fma(uB, CB, rrB); rrB = fma(vB, SB, rrB);
        base += len + 1;
      }
      acc += rrA + rrB;
      p++;  // two points consumed
    } else if (VAR == 3) {
      const double rr = sqrt(rho2), inv = 1.0 / rr;
      const double x = s2 * inv, zx = s0 * inv, zy = s1 * inv;
      double u = 1, v = 0, r = 0;
      int base = 0;
      for (int m = 0; m <= L; m++) {
        if (m > 0) { double a1 = v * zy, un = fma(u, zx, -a1), a2 = v * zx, vn = fma(u, zy, a2); u = un; v = vn; }
        const double2 c0 = s_ab[base];
        double C = c0.x, S = c0.y;
        const int len = L - m;
        if (len >= 1) {
          double q1 = s_Ap[base + 1] * x, q2 = 1.0;
          const double2 c1 = s_ab[base + 1];
          C = fma(c1.x, q1, C); S = fma(c1.y, q1, S);
          int i = 2;
          double na0 = s_Ap[base + 2], na1 = s_Ap[base + 3], na2 = s_Ap[base + 4], na3 = s_Ap[base + 5];
          double2 nc0 = s_ab[base + 2], nc1 = s_ab[base + 3], nc2 = s_ab[base + 4], nc3 = s_ab[base + 5];
          for (; i + 3 <= len; i += 4) {
            const double a0 = na0, a1 = na1, a2 = na2, a3 = na3; const double2 k0 = nc0, k1 = nc1, k2 = nc2, k3 = nc3;
            na0 = s_Ap[base + i + 4]; na1 = s_Ap[base + i + 5]; na2 = s_Ap[base + i + 6]; na3 = s_Ap[base + i + 7];
            nc0 = s_ab[base + i + 4]; nc1 = s_ab[base + i + 5]; nc2 = s_ab[base + i + 6]; nc3 = s_ab[base + i + 7];
            double q;
            q = fma(a0 * x, q1, -q2); C = fma(k0.x, q, C); S = fma(k0.y, q, S); q2 = q1; q1 = q;
            q = fma(a1 * x, q1, -q2); C = fma(k1.x, q, C); S = fma(k1.y, q, S); q2 = q1; q1 = q;
            q = fma(a2 * x, q1, -q2); C = fma(k2.x, q, C); S = fma(k2.y, q, S); q2 = q1; q1 = q;
            q = fma(a3 * x, q1, -q2); C = fma(k3.x, q, C); S = fma(k3.y, q, S); q2 = q1; q1 = q;
          }
          // remainder 0..3 terms from the prefetched registers
          if (i <= len) { double q = fma(na0 * x, q1, -q2); C = fma(nc0.x, q, C); S = fma(nc0.y, q, S); q2 = q1; q1 = q; i++; }
          if (i <= len) { double q = fma(na1 * x, q1, -q2); C = fma(nc1.x, q, C); S = fma(nc1.y, q, S); q2 = q1; q1 = q; i++; }
          if (i <= len) { double q = fma(na2 * x, q1, -q2); C = fma(nc2.x, q, C); S = fma(nc2.y, q, S); q2 = q1; q1 = q; i++; }
        }
        r = fma(u, C, r); r = fma(v, S, r);
        base += len + 1;
      }
      acc += r;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int nsm = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %.0f MHz\n", prop.name, nsm, prop.clockRate / 1e3);
  double *out;
  CK(cudaMalloc(&out, sizeof(double) * nsm * 64 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double clk = 1.965e9;
  // ---- 1. DFMA chains
  auto run_chain = [&](int ilp, int warps_per_sm) {
    const int threads = 128, blocks = nsm * warps_per_sm / 4, iters = 2048;
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0));
      if (ilp == 1) dfma_chain<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      else if (ilp == 2) dfma_chain<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      else if (ilp == 4) dfma_chain<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      else dfma_chain<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    const double n_dfma_per_warp = (double)iters * 16 * ilp;
    const double cyc = best * 1e-3 * clk;
    const double wps = warps_per_sm / 4.0;  // warps per SMSP
    printf("chain ilp %d warps/SM %2d: %.2f cycles per DFMA per warp; pipe util (2 cyc/instr) %.1f%%\n", ilp, warps_per_sm,
           cyc / n_dfma_per_warp, 100.0 * n_dfma_per_warp * wps * 2.0 / cyc);
  };
  for (int ilp : {1, 2, 4, 8}) for (int w : {4, 8, 16}) run_chain(ilp, w);
  // ---- 2. inner loop variants
  for (int L : {30, 20}) {
    const int T = (L + 1) * (L + 2) / 2, T4 = (T + 3) / 4 * 4;
    std::vector<double> Ap(T4 + 8, 0.0);
    std::vector<double2> ab(T4 + 8, make_double2(0, 0));
    int o = 0;
    for (int m = 0; m <= L; m++) { for (int l = m; l <= L; l++) { Ap[o] = (l == m) ? 0.0 : 1.9 + 0.001 * l; ab[o] = make_double2(1e-3 / (1 + l), 2e-3 / (1 + l + m)); o++; } }
    double *dAp; double2 *dab;
    CK(cudaMalloc(&dAp, Ap.size() * 8)); CK(cudaMalloc(&dab, ab.size() * 16));
    CK(cudaMemcpy(dAp, Ap.data(), Ap.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dab, ab.data(), ab.size() * 16, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(T4 + 8) * 24 + 64;
    for (int var = 0; var < 5; var++) {
      for (int warps : {8, 12, 16, 24, 32}) {
        const int threads = 128;
        const int blocks_per_sm = warps / 4;
        const int blocks = nsm * blocks_per_sm, npts = 16;
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
          CK(cudaEventRecord(e0));
          switch (var) {
            case 0: eval_loop<0><<<blocks, threads, smem>>>(out, L, T4, dAp, dab, npts); break;
            case 1: eval_loop<1><<<blocks, threads, smem>>>(out, L, T4, dAp, dab, npts); break;
            case 2: eval_loop<2><<<blocks, threads, smem>>>(out, L, T4, dAp, dab, npts); break;
            case 3: eval_loop<3><<<blocks, threads, smem>>>(out, L, T4, dAp, dab, npts); break;
            case 4: eval_loop<4><<<blocks, threads, smem>>>(out, L, T4, dAp, dab, npts); break;
          }
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
        }
        CK(cudaGetLastError());
        const double slots_per_pt = 4.0 * T + 10.0 * (L + 1) + 30;  // SURVEY I_eval model
        const double cyc = best * 1e-3 * clk;
        const double util = npts * slots_per_pt * (warps / 4.0) * 2.0 / cyc;
        printf("L %d var %d warps/SM %2d: %.3f ms, model pipe util %.1f%%  (%.1f ns per point-eval per warp)\n", L, var, warps, best,
               100 * util, best * 1e6 / npts);
      }
    }
    cudaFree(dAp); cudaFree(dab);
  }
  return 0;
}
