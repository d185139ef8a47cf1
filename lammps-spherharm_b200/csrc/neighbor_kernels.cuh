// neighbor_kernels.cuh — GPU bounding-sphere binned neighbor build (K1-K2).
// Replaces the Neighbor/NBin/NPair path the reference's pair style requests its list from
// (upstream LAMMPS; reference source NOT IN MOUNT).  Output:
//   full list  : CSR nbr_off[n+1], nbr_j[e]      (every neighbor of every atom; gather order)
//   pair list  : pair_i/pair_j (i<j), pair_eij, pair_eji = the two CSR entries of the pair
// All orders are a deterministic function of the atom order (cells sorted by atom index).
#pragma once
#include <cuda_runtime.h>
#include "device_math.cuh"

namespace shgpu {

struct BinGrid {
  double lo[3], len[3], boxlen[3];
  int nc[3], periodic[3];
  double skin;
};

// ---------------- exclusive scan (3-phase) ----------------
constexpr int SCAN_THREADS = 256, SCAN_PER_THREAD = 4, SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__global__ void scan_tile_kernel(const int *in, int *out, int *tile_sum, int n) {
  __shared__ int s_warp[SCAN_THREADS / 32];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
  int v[SCAN_PER_THREAD], tsum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) { v[k] = (base + k < n) ? in[base + k] : 0; tsum += v[k]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < SCAN_THREADS / 32) ? s_warp[lane] : 0, winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
    if (lane < SCAN_THREADS / 32) s_warp[lane] = winc - w;
    if (lane == SCAN_THREADS / 32 - 1) tile_sum[blockIdx.x] = winc;
  }
  __syncthreads();
  int run = s_warp[warp] + inc - tsum;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
}
__global__ void scan_sums_kernel(int *tile_sum, int ntiles, int *total) {
  // single block: serial chunks per thread + block scan
  __shared__ int s_part[1024];
  const int per = (ntiles + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * per, e = min(b + per, ntiles);
  int s = 0;
  for (int k = b; k < e; k++) s += tile_sum[k];
  s_part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 0; k < (int)blockDim.x; k++) { int t = s_part[k]; s_part[k] = run; run += t; }
    *total = run;
  }
  __syncthreads();
  int run = s_part[threadIdx.x];
  for (int k = b; k < e; k++) { int t = tile_sum[k]; tile_sum[k] = run; run += t; }
}
__global__ void scan_add_kernel(int *out, const int *tile_sum, int n, const int *total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += tile_sum[i / SCAN_TILE];
  if (i == 0) out[n] = *total;
}

// ---------------- bounding box of SH origins (non-periodic dims) ----------------
__global__ void bbox_kernel(const double *c, int n, int stride, double *out /*6: min3,max3 as ordered ints*/) {
  // block reduce then atomics on ordered-integer encodings (deterministic: min/max are exact)
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
    for (int d = 0; d < 3; d++) { const double v = c[d * stride + i]; mn[d] = fmin(mn[d], v); mx[d] = fmax(mx[d], v); }
#pragma unroll
  for (int d = 0; d < 3; d++)
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  if ((threadIdx.x & 31) == 0) {
    auto enc = [](double v) { long long b = __double_as_longlong(v); return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL); };
    long long *o = reinterpret_cast<long long *>(out);
#pragma unroll
    for (int d = 0; d < 3; d++) { atomicMin(&o[d], enc(mn[d])); atomicMax(&o[3 + d], enc(mx[d])); }
  }
}

// ---------------- binning ----------------
__device__ __forceinline__ int cell_coord(const BinGrid &G, int d, double v) {
  double u = v - G.lo[d];
  if (G.periodic[d]) u -= G.len[d] * floor(u / G.len[d]);
  int b = (int)(u / G.len[d] * G.nc[d]);
  return min(max(b, 0), G.nc[d] - 1);
}
__global__ void bin_count_kernel(const double *c, int n, int stride, BinGrid G, int *cell_of, int *cell_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cx = cell_coord(G, 0, c[i]), cy = cell_coord(G, 1, c[stride + i]), cz = cell_coord(G, 2, c[2 * stride + i]);
  const int id = (cz * G.nc[1] + cy) * G.nc[0] + cx;
  cell_of[i] = id;
  atomicAdd(&cell_count[id], 1);
}
__global__ void bin_fill_kernel(int n, const int *cell_of, const int *cell_start, int *cell_fill, int *cell_atoms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = cell_of[i];
  cell_atoms[cell_start[id] + atomicAdd(&cell_fill[id], 1)] = i;
}
__global__ void bin_sort_kernel(int ncell, const int *cell_start, int *cell_atoms) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= ncell) return;
  const int b = cell_start[id], e = cell_start[id + 1];
  for (int p = b + 1; p < e; p++) {
    const int key = cell_atoms[p];
    int q = p - 1;
    while (q >= b && cell_atoms[q] > key) { cell_atoms[q + 1] = cell_atoms[q]; q--; }
    cell_atoms[q + 1] = key;
  }
}

// visit every candidate j of atom i in the 27-cell stencil, in a fixed order
template <class F>
__device__ __forceinline__ void for_each_neighbor(const double *c, const int *shape, const DevShape *shapes, int stride,
                                                  const BinGrid &G, const int *cell_of, const int *cell_start,
                                                  const int *cell_atoms, int i, F &&fn) {
  const int id = cell_of[i];
  const int cx = id % G.nc[0], cy = (id / G.nc[0]) % G.nc[1], cz = id / (G.nc[0] * G.nc[1]);
  const double ci0 = c[i], ci1 = c[stride + i], ci2 = c[2 * stride + i];
  const double ri = shapes[shape[i]].rmax;
  int lo[3], hi[3];
  const int cc[3] = {cx, cy, cz};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (G.nc[d] == 1) { lo[d] = hi[d] = 0; }
    else if (G.periodic[d]) { lo[d] = cc[d] - 1; hi[d] = cc[d] + 1; }
    else { lo[d] = max(cc[d] - 1, 0); hi[d] = min(cc[d] + 1, G.nc[d] - 1); }
  }
  for (int z = lo[2]; z <= hi[2]; z++)
    for (int y = lo[1]; y <= hi[1]; y++)
      for (int x = lo[0]; x <= hi[0]; x++) {
        const int wx = (x + G.nc[0]) % G.nc[0], wy = (y + G.nc[1]) % G.nc[1], wz = (z + G.nc[2]) % G.nc[2];
        const int cid = (wz * G.nc[1] + wy) * G.nc[0] + wx;
        for (int q = cell_start[cid]; q < cell_start[cid + 1]; q++) {
          const int j = cell_atoms[q];
          if (j == i) continue;
          double d0 = ci0 - c[j], d1 = ci1 - c[stride + j], d2 = ci2 - c[2 * stride + j];
          // periodic image of the pair: n = rint(d / L) in {-1, 0, 1}, packed 2 bits per dimension (n + 1).  Neighbors are
          // closer than half a box, so n cannot change before the next build and the pair kernels reuse it.
          int img = 1 | (1 << 2) | (1 << 4);
          if (G.periodic[0]) { const double n = rint(d0 / G.boxlen[0]); d0 = d0 - G.boxlen[0] * n; img += (int)n; }
          if (G.periodic[1]) { const double n = rint(d1 / G.boxlen[1]); d1 = d1 - G.boxlen[1] * n; img += (int)n << 2; }
          if (G.periodic[2]) { const double n = rint(d2 / G.boxlen[2]); d2 = d2 - G.boxlen[2] * n; img += (int)n << 4; }
          const double rc = ri + shapes[shape[j]].rmax + G.skin;
          if (d0 * d0 + d1 * d1 + d2 * d2 < rc * rc) fn(j, img);
        }
      }
}

// Which atoms are neighbours of row a.  newton off (tags == nullptr): rows exist for owned atoms only and every
// neighbour counts (a pair that straddles a rank boundary is evaluated by both ranks, each keeps its own half).
// newton on: a pair with a ghost is evaluated by exactly ONE rank, chosen from the two tags (the owner of the smaller tag
// when their sum is even, of the larger one when odd: balanced and symmetric); the ghost then needs a row of its own to
// collect the reaction, which is sent back to its owner (reverse communication).
__device__ __forceinline__ bool pair_is_mine(long long tag_owned, long long tag_ghost) {
  const bool owned_is_smaller = tag_owned < tag_ghost;
  return (((tag_owned + tag_ghost) & 1LL) == 0) == owned_is_smaller;
}
__device__ __forceinline__ bool nbr_included(int a, int b, int nown, const long long *tags) {
  if (!tags) return true;                          // newton off: a is owned
  const bool ga = a >= nown, gb = b >= nown;
  if (!ga && !gb) return true;
  if (ga && gb) return false;
  return ga ? pair_is_mine(tags[b], tags[a]) : pair_is_mine(tags[a], tags[b]);
}
__global__ void nbr_count_kernel(const double *c, const int *shape, const DevShape *shapes, int nrows, int nown, const long long *tags,
                                 int stride, BinGrid G, const int *cell_of, const int *cell_start, const int *cell_atoms,
                                 int *cnt_full, int *cnt_half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  int nf = 0, nh = 0;
  for_each_neighbor(c, shape, shapes, stride, G, cell_of, cell_start, cell_atoms, i, [&](int j, int) {
    if (nbr_included(i, j, nown, tags)) { nf++; nh += (i < nown && j > i); }
  });
  cnt_full[i] = nf;
  cnt_half[i] = nh;
}
__global__ void nbr_fill_kernel(const double *c, const int *shape, const DevShape *shapes, int nrows, int nown, const long long *tags,
                                int stride, BinGrid G, const int *cell_of, const int *cell_start, const int *cell_atoms,
                                const int *nbr_off, const int *half_off, int *nbr_j, int *pair_i, int *pair_j,
                                int *pair_eij, int *pair_img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  int e = nbr_off[i], h = half_off[i];
  for_each_neighbor(c, shape, shapes, stride, G, cell_of, cell_start, cell_atoms, i, [&](int j, int img) {
    if (!nbr_included(i, j, nown, tags)) return;
    nbr_j[e] = j;
    if (i < nown && j > i) { pair_i[h] = i; pair_j[h] = j; pair_eij[h] = e; pair_img[h] = img; h++; }
    e++;
  });
}
// reverse CSR entry of every pair.  Only atoms with index < nrows have CSR rows (newton off: the owned atoms; for an
// owned-ghost pair there is then no reverse entry and nbr_off[j] must not be read at all).
__global__ void pair_reverse_kernel(int npairs, int nrows, const int *pair_i, const int *pair_j, const int *nbr_off,
                                    const int *nbr_j, int *pair_eji) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const int i = pair_i[p], j = pair_j[p];
  int found = -1;
  if (j < nrows)
    for (int e = nbr_off[j]; e < nbr_off[j + 1]; e++)
      if (nbr_j[e] == i) { found = e; break; }
  pair_eji[p] = found;
}

__global__ void copy_origin_kernel(const double *c, double *c0, int n, int stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { c0[i] = c[i]; c0[stride + i] = c[stride + i]; c0[2 * stride + i] = c[2 * stride + i]; }
}

}  // namespace shgpu
