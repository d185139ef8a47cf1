// device_math.cuh — FP64 device arithmetic of the SPHERHARM contact path (sm_100a).
//
// Compiled with -fmad=false: the compiler never contracts a*b+c on its own; every fused
// multiply-add below is an explicit fma() (DFMA).  The sequence of operations on the node
// decision path (pose -> relative pose -> node transform -> rho^2 -> folded SH radius) is the
// contract of DESIGN.md §3, shared with (but not included from) the CPU oracle, so that the
// inside/outside decision of every surface node is bit-identical on both sides.
// Reference source (pair_style spherharm, SPHERHARM math namespace): NOT IN MOUNT.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace shgpu {

struct DevShape {
  int lmax, nterms, nq, nchunks;          // nchunks = ceil(nq/32)
  double rmax, rmin, rmax2, rmin2, mass, inv_mass;
  double inertia[3], com[3], Rp[9];       // Rp row-major: principal frame -> shape frame
  const double *Ap;                       // folded recurrence multipliers, m-major
  const double2 *ab;                      // folded (a,b) coefficients, m-major
  const double *px, *py, *pz;             // node points, shape frame
  const double *nx, *ny, *nz;             // oriented area elements n dS, shape frame
  int n_theta, n_phi;                     // node k = row*n_phi + col; phi_col = (col+1/2) 2pi/n_phi
  int tab_off;                            // offset (in terms) of this shape in the concatenated table
  int pad_;
  const float *row_x;                     // cos(theta_row), float copy for the conservative window
};

// rotation matrix (row-major R[3*r+c]) of unit quaternion (w,x,y,z); plain mul/add, fixed order
__device__ __forceinline__ void quat_to_mat(const double q[4], double R[9]) {
  const double w2 = q[0] * q[0], i2 = q[1] * q[1], j2 = q[2] * q[2], k2 = q[3] * q[3];
  const double twoij = (2.0 * q[1]) * q[2], twoik = (2.0 * q[1]) * q[3], twojk = (2.0 * q[2]) * q[3];
  const double twoiw = (2.0 * q[1]) * q[0], twojw = (2.0 * q[2]) * q[0], twokw = (2.0 * q[3]) * q[0];
  R[0] = ((w2 + i2) - j2) - k2;  R[1] = twoij - twokw;          R[2] = twojw + twoik;
  R[3] = twoij + twokw;          R[4] = ((w2 - i2) + j2) - k2;  R[5] = twojk - twoiw;
  R[6] = twoik - twojw;          R[7] = twojk + twoiw;          R[8] = ((w2 - i2) - j2) + k2;
}

// Folded, trig-free spherical-harmonic radius (SURVEY A.2 / App. B "folded form").
// s = point in the shape frame, rho2 = |s|^2 (already computed by the caller with the
// contract's fma chain).  Tables may live in shared or global memory.
// Per (l,m) term: 1 DMUL + 3 DFMA.
__device__ __forceinline__ double sh_radius_folded(int L, const double *__restrict__ Ap,
                                                   const double2 *__restrict__ ab, double s0, double s1,
                                                   double s2, double rho2, double &rho_out) {
  const double rho = sqrt(rho2);
  const double inv = 1.0 / rho;
  const double x = s2 * inv, zx = s0 * inv, zy = s1 * inv;
  double u = 1.0, v = 0.0, r = 0.0;
  int base = 0;
  for (int m = 0; m <= L; m++) {
    if (m > 0) {
      const double t1 = v * zy, un = fma(u, zx, -t1);
      const double t2 = v * zx, vn = fma(u, zy, t2);
      u = un; v = vn;
    }
    const double2 c0 = ab[base];
    double C = c0.x, S = c0.y;
    const int len = L - m;
    if (len >= 1) {
      double q1 = Ap[base + 1] * x, q2 = 1.0;
      const double2 c1 = ab[base + 1];
      C = fma(c1.x, q1, C); S = fma(c1.y, q1, S);
#pragma unroll 4
      for (int i = 2; i <= len; i++) {
        const double tx = Ap[base + i] * x;
        const double q = fma(tx, q1, -q2);
        const double2 ci = ab[base + i];
        C = fma(ci.x, q, C); S = fma(ci.y, q, S);
        q2 = q1; q1 = q;
      }
    }
    r = fma(u, C, r); r = fma(v, S, r);
    base += len + 1;
  }
  rho_out = rho;
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace shgpu
