// decomp_kernels.cuh — device side of the spatial domain decomposition (row a12).
// Replaces the Comm::exchange / Comm::borders / forward_comm path of upstream LAMMPS' CommBrick (reference source NOT
// IN MOUNT) for the SPHERHARM atom style: everything that touches per-atom data runs on the device; the host only sees
// a handful of counters per neighbor rebuild.  Lees-Edwards shear (BASELINE configs[3], "periodic shear box"): flow
// along x, gradient along y; an atom (or ghost image) that crosses the y boundary n times is displaced by -n * le_offset
// along x and its velocity by -n * le_vshear (fix deform xy + remap v in LAMMPS terms).
#pragma once
#include "step_kernels.cuh"

namespace shgpu {

constexpr int DD_MAX_RANKS = 64;

struct DdGeom {
  double glo[3], L[3], mylo[3], myhi[3], sub[3], rc;
  int gper[3], pgrid[3], ghosted[3];
  int nslot;                 // active neighbour offsets, sorted by (destination rank, offset id)
  int off[26][3];
  double shift[26][3];       // added to x of an atom sent through this slot (periodic image; includes the current
                             // Lees-Edwards offset for slots that cross y)
  double vshift[26];         // added to v_x of such an image
  double le_offset, le_vshear;   // current image offset along x per +1 crossing of y, and its rate * L_y
  signed char key_of_rank[DD_MAX_RANKS];   // migration key of a destination rank: 0 = stays, 1 + k = k-th distinct
                                           // neighbour rank, -1 = not a neighbour (lost atom)
};

// wrap owned positions into the global box (periodic dims), find the owning brick of every owned atom and raise its
// key flag: flag[key * n + i] = 1 (the matrix is zeroed by the caller)
__global__ void dd_wrap_owner_kernel(AtomView A, DdGeom G, int *flag, int *lost) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int st = A.stride;
  double x[3] = {A.x[i], A.x[st + i], A.x[2 * st + i]};
  if (G.gper[1]) {   // y first: a crossing shifts x (Lees-Edwards), then x wraps
    const double ny = floor((x[1] - G.glo[1]) / G.L[1]);
    if (ny != 0.0) {
      x[1] -= G.L[1] * ny;
      if (G.le_offset != 0.0 || G.le_vshear != 0.0) { x[0] -= ny * G.le_offset; A.v[i] -= ny * G.le_vshear; }
    }
  }
  if (G.gper[0]) x[0] -= G.L[0] * floor((x[0] - G.glo[0]) / G.L[0]);
  if (G.gper[2]) x[2] -= G.L[2] * floor((x[2] - G.glo[2]) / G.L[2]);
  int gi[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    A.x[d * st + i] = x[d];
    const int c = (int)floor((x[d] - G.glo[d]) / G.sub[d]);
    gi[d] = min(max(c, 0), G.pgrid[d] - 1);
  }
  const int r = (gi[0] * G.pgrid[1] + gi[1]) * G.pgrid[2] + gi[2];
  int key = G.key_of_rank[r];
  if (key < 0) { atomicAdd(lost, 1); key = 0; }
  flag[(size_t)key * A.n + i] = 1;
}

// after the exclusive scan `pos` of a (nkey x n) flag matrix: start of every key block, and the stable order list
__global__ void dd_starts_kernel(const int *pos, int nkey, int n, int *starts /* nkey + 1 */) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= nkey) starts[k] = pos[(size_t)k * n];
}
__global__ void dd_order_kernel(size_t total, int n, const int *flag, const int *pos, int *order, int *order_key) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  if (flag[t]) { const int p = pos[t]; order[p] = (int)(t % n); if (order_key) order_key[p] = (int)(t / n); }
}
// counts per destination rank from the key / slot starts: cnt[k] = starts[hi[k]] - starts[lo[k]]
__global__ void dd_counts_kernel(const int *starts, int ngroups, const int *lo, const int *hi, int *cnt) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < ngroups) cnt[k] = starts[hi[k]] - starts[lo[k]];
}

// migration record: tag, shape, x3, v3, quat4, angmom3 = 15 doubles (+1 pad)
#define DD_MIGREC 16
__global__ void dd_pack_migrants_kernel(AtomView A, const long long *tag, int m, const int *list, double *out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int i = list[k], st = A.stride;
  double *o = out + (size_t)DD_MIGREC * k;
  o[0] = (double)tag[i]; o[1] = (double)A.shape[i]; o[2] = 0.0;
#pragma unroll
  for (int d = 0; d < 3; d++) { o[3 + d] = A.x[d * st + i]; o[6 + d] = A.v[d * st + i]; o[13 + d] = A.L[d * st + i]; }
#pragma unroll
  for (int d = 0; d < 4; d++) o[9 + d] = A.q[d * st + i];
}

// new owned arrays: stayers (old order) followed by the arrivals (source rank order)
struct OwnedArrays { double *x, *v, *q, *L; int *shape; long long *tag; int stride; };
__global__ void dd_compact_kernel(OwnedArrays src, OwnedArrays dst, int nstay, const int *stay_list, int narr, const double *arr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nstay) {
    const int i = stay_list[k];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      dst.x[d * dst.stride + k] = src.x[d * src.stride + i];
      dst.v[d * dst.stride + k] = src.v[d * src.stride + i];
      dst.L[d * dst.stride + k] = src.L[d * src.stride + i];
    }
#pragma unroll
    for (int d = 0; d < 4; d++) dst.q[d * dst.stride + k] = src.q[d * src.stride + i];
    dst.shape[k] = src.shape[i]; dst.tag[k] = src.tag[i];
  } else if (k < nstay + narr) {
    const double *o = arr + (size_t)DD_MIGREC * (k - nstay);
#pragma unroll
    for (int d = 0; d < 3; d++) { dst.x[d * dst.stride + k] = o[3 + d]; dst.v[d * dst.stride + k] = o[6 + d]; dst.L[d * dst.stride + k] = o[13 + d]; }
#pragma unroll
    for (int d = 0; d < 4; d++) dst.q[d * dst.stride + k] = o[9 + d];
    dst.shape[k] = (int)o[1]; dst.tag[k] = (long long)o[0];
  }
}
// same arrays, new stride (capacity growth that keeps the first n atoms)
__global__ void dd_restride_kernel(OwnedArrays src, OwnedArrays dst, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    dst.x[d * dst.stride + k] = src.x[d * src.stride + k];
    dst.v[d * dst.stride + k] = src.v[d * src.stride + k];
    dst.L[d * dst.stride + k] = src.L[d * src.stride + k];
  }
#pragma unroll
  for (int d = 0; d < 4; d++) dst.q[d * dst.stride + k] = src.q[d * src.stride + k];
  dst.shape[k] = src.shape[k]; dst.tag[k] = src.tag[k];
}

// border flags: flag[s * nown + i] = owned atom i lies in the shell that neighbour slot s needs as ghosts
__global__ void dd_border_flag_kernel(AtomView A, DdGeom G, int *flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int st = A.stride;
  int hi[3], lo[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double x = A.x[d * st + i];
    hi[d] = G.ghosted[d] && x >= G.myhi[d] - G.rc;
    lo[d] = G.ghosted[d] && x < G.mylo[d] + G.rc;
  }
  for (int s = 0; s < G.nslot; s++) {
    bool in = true;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (G.off[s][d] == 1) in = in && hi[d];
      else if (G.off[s][d] == -1) in = in && lo[d];
    }
    flag[(size_t)s * A.n + i] = in ? 1 : 0;
  }
}

// destination of the per-step records when the exchange goes through peer memory: record k of the send list (which belongs to
// neighbour nbr_of_slot[slot]) lands in that neighbour's inbox at its offset for this rank
struct PeerPlan {
  int enabled, nnbr;
  double *dst[32];            // neighbour's inbox for this parity
  long long off[32];          // my record offset inside it (records)
  int seg_start[32];          // first send-list index of the neighbour's segment
  int nbr_of_slot[26];
};

// ghost records.  border records (full = 1): tag, shape, then the per-step part; per-step part: x + shift, quat
// [, v + vshift, angmom when with_vel]
__global__ void dd_pack_kernel(AtomView A, const long long *tag, DdGeom G, int m, const int *send_idx, const int *send_slot,
                               int full, int with_vel, double *out, const __grid_constant__ PeerPlan P) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int i = send_idx[k], s = send_slot[k], st = A.stride;
  const int w = (full ? 2 : 0) + 7 + (with_vel ? 6 : 0);
  double *o = out + (size_t)w * k;
  if (P.enabled) {            // fused pack + push: the record is stored straight into the neighbour's memory over NVLink
    const int nb = P.nbr_of_slot[s];
    o = P.dst[nb] + (size_t)w * (size_t)(P.off[nb] + (k - P.seg_start[nb]));
  }
  int b = 0;
  if (full) { o[0] = (double)tag[i]; o[1] = (double)A.shape[i]; b = 2; }
#pragma unroll
  for (int d = 0; d < 3; d++) o[b + d] = A.x[d * st + i] + G.shift[s][d];
#pragma unroll
  for (int d = 0; d < 4; d++) o[b + 3 + d] = A.q[d * st + i];
  if (with_vel) {
#pragma unroll
    for (int d = 0; d < 3; d++) { o[b + 7 + d] = A.v[d * st + i] + (d == 0 ? G.vshift[s] : 0.0); o[b + 10 + d] = A.L[d * st + i]; }
  }
}
__global__ void dd_unpack_kernel(AtomView A, long long *tag, int first, int m, int full, int with_vel, const double *in) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int i = first + k, st = A.stride;
  const int w = (full ? 2 : 0) + 7 + (with_vel ? 6 : 0);
  const double *o = in + (size_t)w * k;
  int b = 0;
  if (full) {
    tag[i] = (long long)o[0]; A.shape[i] = (int)o[1]; b = 2;
    if (!with_vel) {
#pragma unroll
      for (int d = 0; d < 3; d++) { A.v[d * st + i] = 0.0; A.L[d * st + i] = 0.0; }
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) A.x[d * st + i] = o[b + d];
#pragma unroll
  for (int d = 0; d < 4; d++) A.q[d * st + i] = o[b + 3 + d];
  if (with_vel) {
#pragma unroll
    for (int d = 0; d < 3; d++) { A.v[d * st + i] = o[b + 7 + d]; A.L[d * st + i] = o[b + 10 + d]; }
  }
}

// ---- reverse communication (newton on): force / torque collected on ghosts go back to the owners ------------------
// per owned atom: in how many slots it is sent, and (after the scan) its positions in the send list in slot order
__global__ void dd_rev_count_kernel(int nown, int nslot, const int *flag, int *cnt) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nown) return;
  int c = 0;
  for (int s = 0; s < nslot; s++) c += flag[(size_t)s * nown + a];
  cnt[a] = c;
}
__global__ void dd_rev_fill_kernel(int nown, int nslot, const int *flag, const int *pos, const int *rev_off, int *rev_k) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nown) return;
  int e = rev_off[a];
  for (int s = 0; s < nslot; s++) if (flag[(size_t)s * nown + a]) rev_k[e++] = pos[(size_t)s * nown + a];
}
__global__ void dd_pack_ghost_forces_kernel(AtomView A, int first, int m, double *out, const __grid_constant__ PeerPlan P) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int i = first + k, st = A.stride;
  double *o = out + 6 * (size_t)k;
  if (P.enabled) {            // ghost k came from neighbour nb (ghosts are ordered by source): return it to that rank's R inbox
    int nb = 0;
    while (nb + 1 < P.nnbr && k >= P.seg_start[nb + 1]) nb++;
    o = P.dst[nb] + 6 * (size_t)(P.off[nb] + (k - P.seg_start[nb]));
  }
#pragma unroll
  for (int d = 0; d < 3; d++) { o[d] = A.f[d * st + i]; o[3 + d] = A.tq[d * st + i]; }
}
// publish "my push number `seq` is complete" to every neighbour, then wait until every neighbour has published the same.
// Launched after the pack kernel on the same stream: its stores are complete (and visible system-wide) when this starts.
// flag rows are indexed by the WRITER's rank.  A wait that lasts longer than ~10 s raises *err instead of hanging the GPU.
struct PeerSync { int nnbr, myrank; unsigned long long *peer_flags[32]; int src_rank[32]; };
__global__ void dd_peer_sync_kernel(const __grid_constant__ PeerSync S, unsigned long long *my_flags, unsigned long long seq, int *err) {
  const int k = threadIdx.x;
  if (k >= S.nnbr) return;
  __threadfence_system();
  *(volatile unsigned long long *)&S.peer_flags[k][S.myrank] = seq;
  __threadfence_system();
  volatile unsigned long long *f = my_flags + S.src_rank[k];
  const long long t0 = clock64();
  while (*f < seq) {
    // ~10 s at 2 GHz; once a wait has failed no later one spins again (the run ends with an error, not with a long hang)
    if (*(volatile int *)err != 0) break;
    if (clock64() - t0 > 20000000000LL) { atomicExch(err, 1); break; }
    __nanosleep(100);
  }
  __threadfence_system();
}
// fixed order per atom (ascending slot): bitwise reproducible
__global__ void dd_add_returned_forces_kernel(AtomView A, const int *rev_off, const int *rev_k, const double *in) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.n) return;
  const int e0 = rev_off[a], e1 = rev_off[a + 1];
  if (e0 == e1) return;
  const int st = A.stride;
  double acc[6];
#pragma unroll
  for (int d = 0; d < 3; d++) { acc[d] = A.f[d * st + a]; acc[3 + d] = A.tq[d * st + a]; }
  for (int e = e0; e < e1; e++) {
    const double *r = in + 6 * (size_t)rev_k[e];
#pragma unroll
    for (int d = 0; d < 6; d++) acc[d] += r[d];
  }
#pragma unroll
  for (int d = 0; d < 3; d++) { A.f[d * st + a] = acc[d]; A.tq[d * st + a] = acc[3 + d]; }
}

// ---- candidate-cache carry-over across a decomposed rebuild ---------------------------------------------------------
__device__ __forceinline__ unsigned dd_hash_tag(long long t) {
  unsigned long long z = (unsigned long long)t * 0x9E3779B97F4A7C15ull;
  return (unsigned)(z >> 32);
}
// open-addressing table of the OLD ghosts (entries = old atom index, -1 = empty), keyed by tag; several images of one
// atom may be present
__global__ void dd_ghost_hash_kernel(const long long *old_tag, int nown, int nghost, int *table, int hs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nghost) return;
  const int e = nown + k;
  unsigned slot = dd_hash_tag(old_tag[e]) & (unsigned)(hs - 1);
  while (atomicCAS(&table[slot], -1, e) != -1) slot = (slot + 1) & (unsigned)(hs - 1);
}
// new atom -> old atom: owned stayers through the compaction order, arrivals are new (-1), ghosts through the tag hash +
// proximity of the origin (images of one atom are a box length apart).  The cache's reference state (origin and
// quaternion at the last cache build) follows the atom; a new atom starts from its current state.
__global__ void dd_cache_map_kernel(AtomView A, const long long *tag, int nown, int nstay, const int *order, const long long *old_tag,
                                    const double *old_c, int old_stride, int old_nown, int old_n, const int *table, int hs, double near2,
                                    const double *old_cc0, const double *old_cq0, int *amap, double Ly, double le_doff, double pc0, double pc1,
                                    double pc2, double *acc /* CACHE_FIT_N sums over the matched atoms */) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = a < A.n;
  const int st = A.stride;
  int o = -1;
  if (!valid) {}
  else if (a < nown) { if (a < nstay) o = order[a]; }
  else {
    const long long t = tag[a];
    unsigned slot = dd_hash_tag(t) & (unsigned)(hs - 1);
    for (int probe = 0; probe < hs; probe++) {
      const int e = table[slot];
      if (e < 0) break;
      if (old_tag[e] == t) {
        const double d0 = A.c[a] - old_c[e], d1 = A.c[st + a] - old_c[old_stride + e], d2 = A.c[2 * st + a] - old_c[2 * old_stride + e];
        if (d0 * d0 + d1 * d1 + d2 * d2 < near2) { o = e; break; }
      }
      slot = (slot + 1) & (unsigned)(hs - 1);
    }
  }
  if (valid) amap[a] = o;
  if (o >= 0) {
    // an owned atom that the migration pass wrapped through a periodic boundary (or sheared: Lees-Edwards) jumped by a
    // box vector since the snapshot, which was taken from the same positions before the wrap: the reference origin jumps
    // with it, or the cache would read the jump as a displacement and be rebuilt at every neighbor rebuild of a flow
#pragma unroll
    // Under Lees-Edwards shear an atom that crossed y continues as its own image, whose reference position also carries
    // the image offset AT THE TIME THE CACHE WAS BUILT: x gains ny * (offset now - offset then) on top of the jump.
    double jump[3] = {0, 0, 0};
    if (a < nown) {
#pragma unroll
      for (int d = 0; d < 3; d++) jump[d] = A.c[d * st + a] - old_c[d * old_stride + o];
      if (le_doff != 0.0) jump[0] += rint(-jump[1] / Ly) * le_doff;
    }
#pragma unroll
    for (int d = 0; d < 3; d++) A.cc0[d * st + a] = old_cc0[d * old_stride + o] + jump[d];
#pragma unroll
    for (int d = 0; d < 4; d++) A.cq0[d * st + a] = old_cq0[d * old_stride + o];
  }
  // affine displacement field of the atoms that were already here (the bulk motion / shear since the cache was built)
  double p[3] = {0, 0, 0}, dd[3] = {0, 0, 0};
  if (o >= 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) { p[d] = A.cc0[d * st + a]; dd[d] = A.c[d * st + a] - p[d]; }
  }
  const double pc[3] = {pc0, pc1, pc2};
  cache_fit_accumulate(o >= 0, p, dd, pc, acc);
}
// an atom that is new on this rank starts from its current state, placed where the bulk motion would have taken it from:
// its reference origin is the current one minus the affine displacement field of the others at its position
// (cache_check_kernel measures every atom against that field), its reference quaternion the current one
__global__ void dd_cache_new_atoms_kernel(AtomView A, const int *amap, const double *fit, double pc0, double pc1, double pc2) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A.n || amap[a] >= 0) return;
  const int st = A.stride;
  const double q[3] = {(A.c[a] - pc0) - fit[3], (A.c[st + a] - pc1) - fit[4], (A.c[2 * st + a] - pc2) - fit[5]};
#pragma unroll
  for (int d = 0; d < 3; d++)
    A.cc0[d * st + a] = A.c[d * st + a] - (fit[d] + (fit[6 + 3 * d] * q[0] + fit[7 + 3 * d] * q[1] + fit[8 + 3 * d] * q[2]));
#pragma unroll
  for (int d = 0; d < 4; d++) A.cq0[d * st + a] = A.q[d * st + a];
}

}  // namespace shgpu
