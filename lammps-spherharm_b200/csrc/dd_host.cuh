// dd_host.cuh — host driver of the in-library spatial domain decomposition (SURVEY §8b "sh_create(h, ngpu, ...) with NCCL
// inside", §8e; row a12).  Included by shgpu_api.cu after sh_ctx and its helpers.  One handle = one rank = one GPU; the
// ranks of a job may be processes (torchrun, MPI) or threads of one process (shlmp -gpus N).  Transport: NCCL point-to-point
// (grouped ncclSend/ncclRecv over NVLink) issued on the library's stream; NCCL is bound at run time with dlopen so that a
// single-GPU build of the host code has no NCCL dependency.  All per-atom work (ownership, migration, border lists, ghost
// packing) is done by the kernels in decomp_kernels.cuh; the host reads back ~30 counters per neighbor rebuild.
#pragma once
#include <dlfcn.h>
#include <unistd.h>

namespace {

NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
  if (!api.lib) { api.err = "NCCL not found (dlopen libnccl.so.2)"; return nullptr; }
#define SHGPU_NCCL_SYM(field, name) *(void **)(&api.field) = dlsym(api.lib, name); if (!api.field) { api.err = std::string("NCCL symbol missing: ") + name; api.lib = nullptr; return nullptr; }
  SHGPU_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") SHGPU_NCCL_SYM(CommInitRank, "ncclCommInitRank") SHGPU_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  SHGPU_NCCL_SYM(Send, "ncclSend") SHGPU_NCCL_SYM(Recv, "ncclRecv") SHGPU_NCCL_SYM(AllReduce, "ncclAllReduce") SHGPU_NCCL_SYM(AllGather, "ncclAllGather")
  SHGPU_NCCL_SYM(GroupStart, "ncclGroupStart") SHGPU_NCCL_SYM(GroupEnd, "ncclGroupEnd") SHGPU_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SHGPU_NCCL_SYM
  return &api;
}

#define NC(call) do { ncclResult_t _r = (call); if (_r != ncclSuccess) return fail(h, std::string(#call) + ": " + h->dd.nccl->GetErrorString(_r)); } while (0)

// factor nranks into (px,py,pz) minimising the ghost surface per rank; dims listed in `forbid` stay undivided
void dd_proc_grid(int nranks, const double L[3], const int forbid[3], int out[3]) {
  double best = -1;
  out[0] = 1; out[1] = 1; out[2] = nranks;
  for (int px = 1; px <= nranks; px++) {
    if (nranks % px || (forbid[0] && px > 1)) continue;
    for (int py = 1; py <= nranks / px; py++) {
      if ((nranks / px) % py || (forbid[1] && py > 1)) continue;
      const int pz = nranks / px / py;
      if (forbid[2] && pz > 1) continue;
      const double sx = L[0] / px, sy = L[1] / py, sz = L[2] / pz;
      double cost = sx * sy * (pz > 1) + sy * sz * (px > 1) + sx * sz * (py > 1);
      cost *= 1.0 + 1e-3 * (sx * sx + sy * sy + sz * sz) / (sx * sy + sy * sz + sx * sz);
      cost += 1e-9 * (px * 100 + py * 10 + pz);
      if (best < 0 || cost < best) { best = cost; out[0] = px; out[1] = py; out[2] = pz; }
    }
  }
}

// current Lees-Edwards image offset along x (per +1 crossing of y).  It is wrapped into [-Lx/2, Lx/2) only at neighbor
// rebuilds: between rebuilds the images must move continuously (the pair images stored at the build stay valid).
double le_offset_now(const sh_ctx *h) { return h->dd.le_off_build + h->dd.le_rate * h->dd.glen[1] * (h->time - h->dd.le_time_build); }

// geometry of this rank's brick, its neighbour slots and their periodic shifts
int dd_setup_geometry(sh_ctx *h) {
  DdCtx &D = h->dd;
  if (!h->box_set) return fail(h, "domain decomposition needs sh_set_box (the GLOBAL box)");
  double rmaxg = 0, commax = 0;
  for (auto &s : h->shapes) {
    rmaxg = std::max(rmaxg, s.rmax);
    commax = std::max(commax, std::sqrt(s.com[0] * s.com[0] + s.com[1] * s.com[1] + s.com[2] * s.com[2]));
  }
  DdGeom &G = D.G;
  G.rc = 2.0 * rmaxg + h->skin + 2.0 * commax;
  const bool le = D.le_rate != 0.0;
  for (int d = 0; d < 3; d++) { D.glen[d] = D.ghi[d] - D.glo[d]; }
  if (!D.pgrid_set) {
    int forbid[3] = {le ? 1 : 0, 0, 0};   // the shear direction stays undivided: images slide along x
    dd_proc_grid(D.nranks, D.glen, forbid, D.pgrid);
  }
  if (D.pgrid[0] * D.pgrid[1] * D.pgrid[2] != D.nranks) return fail(h, "processor grid does not match the number of ranks");
  if (le && D.pgrid[0] > 1) return fail(h, "Lees-Edwards shear: the flow direction x must not be decomposed");
  if (le && !(D.gper[0] && D.gper[1])) return fail(h, "Lees-Edwards shear needs a box periodic in x and y");
  D.g[0] = D.rank / (D.pgrid[1] * D.pgrid[2]); D.g[1] = (D.rank / D.pgrid[2]) % D.pgrid[1]; D.g[2] = D.rank % D.pgrid[2];
  for (int d = 0; d < 3; d++) {
    G.glo[d] = D.glo[d]; G.L[d] = D.glen[d]; G.gper[d] = D.gper[d]; G.pgrid[d] = D.pgrid[d];
    G.sub[d] = D.glen[d] / D.pgrid[d];
    G.mylo[d] = D.glo[d] + D.g[d] * G.sub[d]; G.myhi[d] = G.mylo[d] + G.sub[d];
    G.ghosted[d] = D.pgrid[d] > 1 || (d == 1 && le) || (D.self_ghosts && D.gper[d]);
    if (G.ghosted[d] && G.sub[d] < G.rc * (D.pgrid[d] == 1 ? 2.0 : 1.0)) return fail(h, "sub-domain thinner than the ghost cutoff");
    // the engine sees a dimension as periodic only when it is neither decomposed nor sheared
    h->periodic[d] = D.gper[d] && !G.ghosted[d];
    h->lo[d] = D.glo[d]; h->hi[d] = D.ghi[d];
  }
  // neighbour slots
  struct Slot { int rank, id, off[3]; double shift[3]; };
  std::vector<Slot> slots;
  int id = 0;
  for (int ox = -1; ox <= 1; ox++) for (int oy = -1; oy <= 1; oy++) for (int oz = -1; oz <= 1; oz++, id++) {
    const int o[3] = {ox, oy, oz};
    if (!ox && !oy && !oz) continue;
    int gg[3]; double sh[3] = {0, 0, 0};
    bool ok = true;
    for (int d = 0; d < 3 && ok; d++) {
      gg[d] = D.g[d];
      if (o[d] == 0) continue;
      if (!G.ghosted[d]) { ok = false; break; }
      gg[d] += o[d];
      if (gg[d] >= D.pgrid[d]) { if (!D.gper[d]) ok = false; else { gg[d] -= D.pgrid[d]; sh[d] = -D.glen[d]; } }
      else if (gg[d] < 0) { if (!D.gper[d]) ok = false; else { gg[d] += D.pgrid[d]; sh[d] = D.glen[d]; } }
    }
    if (!ok) continue;
    Slot s; s.rank = (gg[0] * D.pgrid[1] + gg[1]) * D.pgrid[2] + gg[2]; s.id = id;
    for (int d = 0; d < 3; d++) { s.off[d] = o[d]; s.shift[d] = sh[d]; }
    slots.push_back(s);
  }
  std::stable_sort(slots.begin(), slots.end(), [](const Slot &a, const Slot &b) { return a.rank != b.rank ? a.rank < b.rank : a.id < b.id; });
  G.nslot = (int)slots.size();
  D.nbr_rank.clear(); D.slot_lo.clear(); D.slot_hi.clear();
  D.base_shift.assign(26 * 3, 0.0);
  for (int s = 0; s < G.nslot; s++) {
    for (int d = 0; d < 3; d++) { G.off[s][d] = slots[s].off[d]; D.base_shift[3 * s + d] = slots[s].shift[d]; }
    if (D.nbr_rank.empty() || D.nbr_rank.back() != slots[s].rank) { D.nbr_rank.push_back(slots[s].rank); D.slot_lo.push_back(s); D.slot_hi.push_back(s + 1); }
    else D.slot_hi.back() = s + 1;
  }
  // migration keys: distinct neighbour ranks other than this one
  for (int r = 0; r < DD_MAX_RANKS; r++) G.key_of_rank[r] = -1;
  G.key_of_rank[D.rank] = 0;
  D.mig_rank.clear();
  for (int r : D.nbr_rank) if (r != D.rank) { D.mig_rank.push_back(r); G.key_of_rank[r] = (signed char)D.mig_rank.size(); }
  D.geometry_ok = true;
  return 0;
}

// slot shifts for the current time (the Lees-Edwards offset slides)
void dd_refresh_shifts(sh_ctx *h) {
  DdCtx &D = h->dd;
  DdGeom &G = D.G;
  const double off = D.le_rate != 0.0 ? le_offset_now(h) : 0.0, vs = D.le_rate * D.glen[1];
  G.le_offset = off; G.le_vshear = vs;
  for (int s = 0; s < G.nslot; s++) {
    const double nimg = D.base_shift[3 * s + 1] == 0.0 ? 0.0 : (D.base_shift[3 * s + 1] > 0 ? 1.0 : -1.0);
    G.shift[s][0] = D.base_shift[3 * s] + nimg * off;
    G.shift[s][1] = D.base_shift[3 * s + 1];
    G.shift[s][2] = D.base_shift[3 * s + 2];
    G.vshift[s] = nimg * vs;
  }
}

OwnedArrays dd_owned(sh_ctx *h, int which) {
  OwnedArrays O;
  if (which == 0) { O.x = h->x.p; O.v = h->v.p; O.q = h->q.p; O.L = h->L.p; O.shape = h->shape.p; O.tag = h->d_tag.p; O.stride = h->stride; }
  else { O.x = h->dd.x2.p; O.v = h->dd.v2.p; O.q = h->dd.q2.p; O.L = h->dd.L2.p; O.shape = h->dd.shape2.p; O.tag = h->dd.tag2.p; O.stride = h->dd.stride2; }
  return O;
}

// (re)allocate the per-atom arrays that are derived every step for a new stride
int dd_ensure_derived(sh_ctx *h, int st) {
  try {
    h->f.ensure(3 * (size_t)st); h->tq.ensure(3 * (size_t)st); h->c.ensure(3 * (size_t)st); h->Rs.ensure(9 * (size_t)st);
    h->c0.ensure(3 * (size_t)st); h->wallf.ensure(6 * (size_t)st); h->ewall.ensure(st); h->ke.ensure(2 * (size_t)st);
    h->cc0.ensure(3 * (size_t)st); h->cq0.ensure(4 * (size_t)st); h->gf.ensure(6 * (size_t)st);
  } catch (std::string &e) { return fail(h, e); }
  return 0;
}

int dd_alloc_alt(sh_ctx *h, int st) {
  DdCtx &D = h->dd;
  try {
    D.x2.ensure(3 * (size_t)st); D.v2.ensure(3 * (size_t)st); D.q2.ensure(4 * (size_t)st); D.L2.ensure(3 * (size_t)st);
    D.shape2.ensure(st); D.tag2.ensure(st);
  } catch (std::string &e) { return fail(h, e); }
  D.stride2 = st;
  return 0;
}
void dd_swap_alt(sh_ctx *h) {
  DdCtx &D = h->dd;
  std::swap(h->x, D.x2); std::swap(h->v, D.v2); std::swap(h->q, D.q2); std::swap(h->L, D.L2);
  std::swap(h->shape, D.shape2); std::swap(h->d_tag, D.tag2); std::swap(h->stride, D.stride2);
}

// grouped point-to-point exchange with the distinct neighbour ranks `ranks`: element counts sc / rc (in units of `width`
// doubles), displacements are the running sums.  A neighbour that is this rank itself is a device copy.
int dd_sendrecv(sh_ctx *h, const std::vector<int> &ranks, const double *sbuf, const std::vector<int> &sc, double *rbuf,
                const std::vector<int> &rc, int width) {
  DdCtx &D = h->dd;
  size_t so = 0, ro = 0;
  bool any = false;
  for (size_t k = 0; k < ranks.size(); k++) if (ranks[k] != D.rank && (sc[k] > 0 || rc[k] > 0)) any = true;
  if (any) NC(D.nccl->GroupStart());
  for (size_t k = 0; k < ranks.size(); k++) {
    if (ranks[k] == D.rank) {
      if (sc[k] != rc[k]) return fail(h, "self exchange: send and receive counts differ");
      if (sc[k] > 0) CU(cudaMemcpyAsync(rbuf + ro * width, sbuf + so * width, (size_t)sc[k] * width * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    } else {
      if (sc[k] > 0) NC(D.nccl->Send(sbuf + so * width, (size_t)sc[k] * width, ncclDouble, ranks[k], D.comm, h->stream));
      if (rc[k] > 0) NC(D.nccl->Recv(rbuf + ro * width, (size_t)rc[k] * width, ncclDouble, ranks[k], D.comm, h->stream));
    }
    so += sc[k]; ro += rc[k];
  }
  if (any) NC(D.nccl->GroupEnd());
  return 0;
}
// one int to / from every neighbour rank (device buffers d_s[k], d_r[k])
int dd_exchange_counts(sh_ctx *h, const std::vector<int> &ranks, const int *d_s, int *d_r) {
  DdCtx &D = h->dd;
  bool any = false;
  for (int r : ranks) if (r != D.rank) any = true;
  if (any) NC(D.nccl->GroupStart());
  for (size_t k = 0; k < ranks.size(); k++) {
    if (ranks[k] == D.rank) CU(cudaMemcpyAsync(d_r + k, d_s + k, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    else { NC(D.nccl->Send(d_s + k, 1, ncclInt, ranks[k], D.comm, h->stream)); NC(D.nccl->Recv(d_r + k, 1, ncclInt, ranks[k], D.comm, h->stream)); }
  }
  if (any) NC(D.nccl->GroupEnd());
  return 0;
}

// ---- NVLink peer-memory inboxes (PeerCtx in dd_types.h) --------------------------------------------------------------
struct PeerBlob { cudaIpcMemHandle_t handle; long long pid, ptr, capF, capR; int device, ok; char pad[128 - 64 - 4 * 8 - 2 * 4]; };
static_assert(sizeof(PeerBlob) == 128, "PeerBlob is exchanged as 128 bytes");

void dd_peer_release(sh_ctx *h) {
  PeerCtx &P = h->dd.peer;
  for (void *q : P.opened) cudaIpcCloseMemHandle(q);
  P.opened.clear(); P.base.clear();
  if (P.region) { cudaFree(P.region); P.region = nullptr; }
  P.enabled = false;
}

// COLLECTIVE over all ranks: (re)allocate this rank's inbox region for at least needF / needR records, exchange the IPC handles
// and open the neighbours' regions.  Any rank failing switches the peer path off everywhere (NCCL keeps working).
int dd_peer_setup(sh_ctx *h, long long needF, long long needR) {
  DdCtx &D = h->dd;
  PeerCtx &P = D.peer;
  CU(cudaStreamSynchronize(h->stream));
  if (P.region) {
    // growing an existing region: every rank first lets go of its neighbours' regions, and only when ALL have done so does
    // anybody free its own (an exporter must not free memory that an importer still has mapped)
    for (void *q : P.opened) cudaIpcCloseMemHandle(q);
    P.opened.clear();
    int *d_bar = h->dd.d_int.p + 251;
    CU(cudaMemsetAsync(d_bar, 0, sizeof(int), h->stream));
    NC(D.nccl->AllReduce(d_bar, d_bar, 1, ncclInt, ncclMax, D.comm, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  dd_peer_release(h);
  P.capF = std::max<long long>(4096, 2 * needF); P.capR = std::max<long long>(4096, 2 * needR);
  const size_t nd = (size_t)2 * P.capF * P.wmax + (size_t)2 * P.capR * 6 + 2 * DD_MAX_RANKS;
  PeerBlob mine{};
  mine.ok = 1;
  if (cudaMalloc(&P.region, nd * sizeof(double)) != cudaSuccess) { P.region = nullptr; mine.ok = 0; cudaGetLastError(); }
  if (mine.ok) {
    cudaMemset(P.region, 0, nd * sizeof(double));
    if (cudaIpcGetMemHandle(&mine.handle, P.region) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
  }
  if (!P.d_err) { CU(cudaMalloc(&P.d_err, sizeof(int))); CU(cudaMemset(P.d_err, 0, sizeof(int))); }
  mine.pid = (long long)getpid(); mine.ptr = (long long)(uintptr_t)P.region; mine.capF = P.capF; mine.capR = P.capR; mine.device = h->device;
  DevBuf<char> d_blobs;
  try { d_blobs.ensure((size_t)128 * (D.nranks + 1)); P.d_off.ensure(64); } catch (std::string &e) { return fail(h, e); }
  CU(cudaMemcpy(d_blobs.p + (size_t)128 * D.nranks, &mine, 128, cudaMemcpyHostToDevice));
  NC(D.nccl->AllGather(d_blobs.p + (size_t)128 * D.nranks, d_blobs.p, 128, ncclChar, D.comm, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  std::vector<PeerBlob> all(D.nranks);
  CU(cudaMemcpy(all.data(), d_blobs.p, (size_t)128 * D.nranks, cudaMemcpyDeviceToHost));
  d_blobs.release();
  bool ok = true;
  for (auto &b : all) ok = ok && b.ok;
  const int nnb = (int)D.nbr_rank.size();
  P.base.assign(nnb, nullptr); P.pcapF.assign(nnb, 0); P.pcapR.assign(nnb, 0);
  for (int k = 0; k < nnb && ok; k++) {
    const int r = D.nbr_rank[k];
    const PeerBlob &b = all[r];
    P.pcapF[k] = b.capF; P.pcapR[k] = b.capR;
    if (r == D.rank) { P.base[k] = P.region; continue; }
    if (b.pid == mine.pid) {            // rank threads of one process: plain peer access
      int can = 0;
      cudaDeviceCanAccessPeer(&can, h->device, b.device);
      if (!can) { ok = false; break; }
      cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
      cudaGetLastError();
      P.base[k] = (double *)(uintptr_t)b.ptr;
    } else {
      void *q = nullptr;
      if (cudaIpcOpenMemHandle(&q, b.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
      P.opened.push_back(q);
      P.base[k] = (double *)q;
    }
  }
  // everybody or nobody
  int *d_ok = h->dd.d_int.p + 250;
  const int myok = ok ? 1 : 0;
  CU(cudaMemcpy(d_ok, &myok, sizeof(int), cudaMemcpyHostToDevice));
  NC(D.nccl->AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, D.comm, h->stream));
  int allok = 0;
  CU(cudaMemcpyAsync(&allok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (!allok) { dd_peer_release(h); P.want = false; return 0; }
  P.enabled = true; P.seq_f = 0; P.seq_r = 0;
  return 0;
}

double *peer_inbox(const PeerCtx &P, double *base, long long capF, long long capR, bool reverse, int parity, int wmax) {
  (void)P;
  return reverse ? base + (size_t)2 * capF * wmax + (size_t)parity * capR * 6 : base + (size_t)parity * capF * wmax;
}
unsigned long long *peer_flags(double *base, long long capF, long long capR, bool reverse, int wmax) {
  return reinterpret_cast<unsigned long long *>(base + (size_t)2 * capF * wmax + (size_t)2 * capR * 6) + (reverse ? DD_MAX_RANKS : 0);
}

int dd_width(const sh_ctx *h, bool full) { return (full ? 2 : 0) + 7 + (h->dd.ghost_vel ? 6 : 0); }

// Comm::exchange: wrap, find owners, move the atoms that left this brick to their new owners (device compaction)
int dd_migrate(sh_ctx *h) {
  DdCtx &D = h->dd;
  const int nown = (int)(h->n - h->nghost);
  const int nmr = (int)D.mig_rank.size(), nkey = 1 + nmr;
  dd_refresh_shifts(h);
  AtomView A = view(h);
  try {
    D.flag.ensure((size_t)nkey * std::max(nown, 1) + 2); D.pos.ensure((size_t)nkey * std::max(nown, 1) + 2); D.order.ensure((size_t)std::max(nown, 1) + 2);
    D.d_int.ensure(256);
  } catch (std::string &e) { return fail(h, e); }
  int *d_starts = D.d_int.p, *d_cnt = D.d_int.p + 64, *d_rcnt = D.d_int.p + 96, *d_lost = D.d_int.p + 128, *d_lo = D.d_int.p + 160, *d_hi = D.d_int.p + 192;
  CU(cudaMemsetAsync(D.d_int.p, 0, 256 * sizeof(int), h->stream));
  if (nown > 0) {
    CU(cudaMemsetAsync(D.flag.p, 0, (size_t)nkey * nown * sizeof(int), h->stream));
    dd_wrap_owner_kernel<<<cdiv(nown, 256), 256, 0, h->stream>>>(A, D.G, D.flag.p, d_lost);
    if (exclusive_scan(h, D.flag.p, D.pos.p, nkey * nown, h->scalars.p)) return -1;
    dd_starts_kernel<<<1, 64, 0, h->stream>>>(D.pos.p, nkey, nown, d_starts);
    dd_order_kernel<<<cdiv((int64_t)nkey * nown, 256), 256, 0, h->stream>>>((size_t)nkey * nown, nown, D.flag.p, D.pos.p, D.order.p, nullptr);
    h->kernel_launches += 3;
  }
  // counts per migration rank = starts[2+k] - starts[1+k]
  {
    std::vector<int> lo(32, 0), hi(32, 0);
    for (int k = 0; k < nmr; k++) { lo[k] = 1 + k; hi[k] = 2 + k; }
    CU(cudaMemcpyAsync(d_lo, lo.data(), 32 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_hi, hi.data(), 32 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (nmr > 0) { dd_counts_kernel<<<1, 32, 0, h->stream>>>(d_starts, nmr, d_lo, d_hi, d_cnt); h->kernel_launches++; }
  }
  int rc2;
  if ((rc2 = dd_exchange_counts(h, D.mig_rank, d_cnt, d_rcnt))) return rc2;
  CU(cudaMemcpyAsync(D.h_int, D.d_int.p, 160 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (D.h_int[128] != 0) return fail(h, "domain decomposition: an atom moved past a neighbouring brick (lost atom)");
  const int nstay = nown > 0 ? D.h_int[1] - D.h_int[0] : 0;
  std::vector<int> sc(nmr), rc(nmr);
  int nmig = 0, narr = 0;
  for (int k = 0; k < nmr; k++) { sc[k] = D.h_int[64 + k]; rc[k] = D.h_int[96 + k]; nmig += sc[k]; narr += rc[k]; }
  if (nown > 0 && nstay + nmig != nown) return fail(h, "domain decomposition: migration counts are inconsistent");
  D.migrated_out += nmig; D.migrated_in += narr; D.last_nstay = nstay;
  try { D.sendbuf.ensure((size_t)DD_MIGREC * std::max(nmig, 1)); D.recvbuf.ensure((size_t)DD_MIGREC * std::max(narr, 1)); }
  catch (std::string &e) { return fail(h, e); }
  if (nmig > 0) {
    dd_pack_migrants_kernel<<<cdiv(nmig, 256), 256, 0, h->stream>>>(A, h->d_tag.p, nmig, D.order.p + nstay, D.sendbuf.p);
    h->kernel_launches++;
  }
  if ((rc2 = dd_sendrecv(h, D.mig_rank, D.sendbuf.p, sc, D.recvbuf.p, rc, DD_MIGREC))) return rc2;
  // compaction into the alternate arrays (headroom for the ghosts that the border pass appends)
  const int nnew = nstay + narr;
  const int want = (int)(((int64_t)nnew + std::max<int64_t>(h->nghost, nnew / 4) + 31) / 32 * 32) + 64;
  const int st2 = std::max(want, h->stride);
  if ((rc2 = dd_alloc_alt(h, st2))) return rc2;
  if (nnew > 0) {
    dd_compact_kernel<<<cdiv(nnew, 256), 256, 0, h->stream>>>(dd_owned(h, 0), dd_owned(h, 1), nstay, D.order.p, narr, D.recvbuf.p);
    h->kernel_launches++;
  }
  dd_swap_alt(h);
  if ((rc2 = dd_ensure_derived(h, h->stride))) return rc2;
  h->n = nnew; h->nghost = 0;
  h->tags_host_valid = false;
  CU(cudaGetLastError());
  return 0;
}

// Comm::borders: send lists per neighbour slot, ghost creation
int dd_borders(sh_ctx *h) {
  DdCtx &D = h->dd;
  const int nown = (int)(h->n - h->nghost);
  h->n = nown; h->nghost = 0;
  DdGeom &G = D.G;
  const int nslot = G.nslot, nnb = (int)D.nbr_rank.size();
  // wrap the Lees-Edwards offset now: every image jumps by whole box lengths only at a rebuild
  if (D.le_rate != 0.0) {
    double off = le_offset_now(h);
    off -= D.glen[0] * std::floor(off / D.glen[0] + 0.5);
    D.le_off_build = off; D.le_time_build = h->time;
  }
  dd_refresh_shifts(h);
  AtomView A = view(h);
  try {
    D.flag.ensure((size_t)std::max(nslot, 1) * std::max(nown, 1) + 2); D.pos.ensure((size_t)std::max(nslot, 1) * std::max(nown, 1) + 2);
    D.d_int.ensure(256);
  } catch (std::string &e) { return fail(h, e); }
  int *d_starts = D.d_int.p, *d_cnt = D.d_int.p + 64, *d_rcnt = D.d_int.p + 96, *d_lo = D.d_int.p + 160, *d_hi = D.d_int.p + 192;
  CU(cudaMemsetAsync(D.d_int.p, 0, 256 * sizeof(int), h->stream));
  if (nown > 0 && nslot > 0) {
    dd_border_flag_kernel<<<cdiv(nown, 256), 256, 0, h->stream>>>(A, G, D.flag.p);
    if (exclusive_scan(h, D.flag.p, D.pos.p, nslot * nown, h->scalars.p)) return -1;
    dd_starts_kernel<<<1, 64, 0, h->stream>>>(D.pos.p, nslot, nown, d_starts);
    h->kernel_launches += 2;
  }
  {
    std::vector<int> lo(32, 0), hi(32, 0);
    for (int k = 0; k < nnb; k++) { lo[k] = D.slot_lo[k]; hi[k] = D.slot_hi[k]; }
    CU(cudaMemcpyAsync(d_lo, lo.data(), 32 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_hi, hi.data(), 32 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (nnb > 0) { dd_counts_kernel<<<1, 32, 0, h->stream>>>(d_starts, nnb, d_lo, d_hi, d_cnt); h->kernel_launches++; }
  }
  int rc2;
  if ((rc2 = dd_exchange_counts(h, D.nbr_rank, d_cnt, d_rcnt))) return rc2;
  CU(cudaMemcpyAsync(D.h_int, D.d_int.p, 160 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  D.send_cnt.assign(nnb, 0); D.recv_cnt.assign(nnb, 0);
  int nsend = 0, nghost = 0;
  for (int k = 0; k < nnb; k++) { D.send_cnt[k] = D.h_int[64 + k]; D.recv_cnt[k] = D.h_int[96 + k]; nsend += D.send_cnt[k]; nghost += D.recv_cnt[k]; }
  D.nsend = nsend;
  // ---- NVLink peer path: inbox capacities (global decision) and this rank's offsets inside its neighbours' inboxes
  if (D.nranks > 1 && D.peer.want) {
    PeerCtx &P = D.peer;
    int *d_need = D.d_int.p + 240;
    const int need[2] = {P.enabled ? (nghost > P.capF || nsend > P.capR ? 1 : 0) : 1, 0};
    CU(cudaMemcpyAsync(d_need, need, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NC(D.nccl->AllReduce(d_need, d_need, 1, ncclInt, ncclMax, D.comm, h->stream));
    CU(cudaMemcpyAsync(D.h_int + 240, d_need, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (D.h_int[240]) { if ((rc2 = dd_peer_setup(h, nghost, nsend))) return rc2; }
    if (P.enabled) {
      // neighbour k writes its ghosts for me at record roff[k] of my F inbox, and returns forces at soff[k] of my R inbox
      long long offs[64] = {0};
      long long ro = 0, so = 0;
      for (int k = 0; k < nnb; k++) { offs[k] = ro; offs[32 + k] = so; ro += D.recv_cnt[k]; so += D.send_cnt[k]; }
      try { D.recvbuf.ensure(256); } catch (std::string &e) { return fail(h, e); }
      long long *d_mine = reinterpret_cast<long long *>(D.recvbuf.p);   // scratch (everything before is synchronised)
      CU(cudaMemcpyAsync(d_mine, offs, sizeof offs, cudaMemcpyHostToDevice, h->stream));
      bool any = false;
      for (int r : D.nbr_rank) if (r != D.rank) any = true;
      if (any) NC(D.nccl->GroupStart());
      for (int k = 0; k < nnb; k++) {
        if (D.nbr_rank[k] == D.rank) {
          CU(cudaMemcpyAsync(P.d_off.p + k, d_mine + k, sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
          CU(cudaMemcpyAsync(P.d_off.p + 32 + k, d_mine + 32 + k, sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
        } else {
          NC(D.nccl->Send(d_mine + k, 1, ncclInt64, D.nbr_rank[k], D.comm, h->stream));
          NC(D.nccl->Send(d_mine + 32 + k, 1, ncclInt64, D.nbr_rank[k], D.comm, h->stream));
          NC(D.nccl->Recv(P.d_off.p + k, 1, ncclInt64, D.nbr_rank[k], D.comm, h->stream));
          NC(D.nccl->Recv(P.d_off.p + 32 + k, 1, ncclInt64, D.nbr_rank[k], D.comm, h->stream));
        }
      }
      if (any) NC(D.nccl->GroupEnd());
      CU(cudaMemcpyAsync(D.h_off, P.d_off.p, 64 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
    }
  }
  // capacity for the ghosts
  if (nown + nghost + 32 > h->stride) {
    const int st2 = (int)(((int64_t)(nown + nghost) * 5 / 4 + 31) / 32 * 32) + 64;
    if ((rc2 = dd_alloc_alt(h, st2))) return rc2;
    if (nown > 0) { dd_restride_kernel<<<cdiv(nown, 256), 256, 0, h->stream>>>(dd_owned(h, 0), dd_owned(h, 1), nown); h->kernel_launches++; }
    dd_swap_alt(h);
    if ((rc2 = dd_ensure_derived(h, h->stride))) return rc2;
    A = view(h);
  }
  const int wf = dd_width(h, true);
  try {
    D.send_idx.ensure((size_t)std::max(nsend, 1)); D.send_slot.ensure((size_t)std::max(nsend, 1));
    D.sendbuf.ensure((size_t)wf * std::max(nsend, 1)); D.recvbuf.ensure((size_t)wf * std::max(nghost, 1));
  } catch (std::string &e) { return fail(h, e); }
  if (nsend > 0) {
    dd_order_kernel<<<cdiv((int64_t)nslot * nown, 256), 256, 0, h->stream>>>((size_t)nslot * nown, nown, D.flag.p, D.pos.p, D.send_idx.p, D.send_slot.p);
    dd_pack_kernel<<<cdiv(nsend, 256), 256, 0, h->stream>>>(A, h->d_tag.p, G, nsend, D.send_idx.p, D.send_slot.p, 1, D.ghost_vel, D.sendbuf.p, PeerPlan{});
    h->kernel_launches += 2;
  }
  if ((rc2 = dd_sendrecv(h, D.nbr_rank, D.sendbuf.p, D.send_cnt, D.recvbuf.p, D.recv_cnt, wf))) return rc2;
  // reverse communication: where in the send list every owned atom appears
  try { D.rev_cnt.ensure((size_t)nown + 2); D.rev_off.ensure((size_t)nown + 2); D.rev_k.ensure((size_t)std::max(nsend, 1)); }
  catch (std::string &e) { return fail(h, e); }
  if (nown > 0) {
    if (nslot > 0) dd_rev_count_kernel<<<cdiv(nown, 256), 256, 0, h->stream>>>(nown, nslot, D.flag.p, D.rev_cnt.p);
    else CU(cudaMemsetAsync(D.rev_cnt.p, 0, (size_t)nown * sizeof(int), h->stream));
    if (exclusive_scan(h, D.rev_cnt.p, D.rev_off.p, nown, h->scalars.p)) return -1;
    if (nslot > 0 && nsend > 0) dd_rev_fill_kernel<<<cdiv(nown, 256), 256, 0, h->stream>>>(nown, nslot, D.flag.p, D.pos.p, D.rev_off.p, D.rev_k.p);
    h->kernel_launches += 2;
  }
  h->n = nown + nghost; h->nghost = nghost;
  if (nghost > 0) {
    AtomView All = view_all(h);
    dd_unpack_kernel<<<cdiv(nghost, 256), 256, 0, h->stream>>>(All, h->d_tag.p, nown, nghost, 1, D.ghost_vel, D.recvbuf.p);
    h->kernel_launches++;
  }
  h->tags_host_valid = false;
  h->forces_valid = false; h->list_valid = false; h->npairs = 0; h->nentries = 0;
  h->atoms_epoch++; h->cache_state = CACHE_INVALID;
  D.borders_ok = true; D.border_builds++;
  CU(cudaGetLastError());
  return 0;
}

// where this rank's records go in every neighbour's inbox (forward: my send list -> their F inbox; reverse: their ghosts
// here, ordered by source -> their R inbox)
void dd_peer_plan(sh_ctx *h, bool reverse, int parity, PeerPlan &plan) {
  DdCtx &D = h->dd;
  PeerCtx &P = D.peer;
  const int nnb = (int)D.nbr_rank.size();
  plan.enabled = 1; plan.nnbr = nnb;
  int start = 0;
  for (int k = 0; k < nnb; k++) {
    plan.dst[k] = peer_inbox(P, P.base[k], P.pcapF[k], P.pcapR[k], reverse, parity, P.wmax);
    plan.off[k] = D.h_off[(reverse ? 32 : 0) + k];
    plan.seg_start[k] = start;
    start += reverse ? D.recv_cnt[k] : D.send_cnt[k];
    for (int s2 = D.slot_lo[k]; s2 < D.slot_hi[k]; s2++) plan.nbr_of_slot[s2] = k;
  }
}
int dd_peer_sync(sh_ctx *h, bool reverse) {
  DdCtx &D = h->dd;
  PeerCtx &P = D.peer;
  PeerSync S{};
  const int nnb = (int)D.nbr_rank.size();
  S.nnbr = nnb; S.myrank = D.rank;
  for (int k = 0; k < nnb; k++) { S.peer_flags[k] = peer_flags(P.base[k], P.pcapF[k], P.pcapR[k], reverse, P.wmax); S.src_rank[k] = D.nbr_rank[k]; }
  dd_peer_sync_kernel<<<1, 32, 0, h->stream>>>(S, peer_flags(P.region, P.capF, P.capR, reverse, P.wmax), reverse ? P.seq_r : P.seq_f, P.d_err);
  h->kernel_launches++;
  return 0;
}

// forward_comm: ghost x / quat (and v, angmom when a velocity-dependent contact model is on) every step; no host sync
int dd_forward(sh_ctx *h) {
  DdCtx &D = h->dd;
  if (!D.peer.enabled && D.nsend == 0 && h->nghost == 0) return 0;   // (peer path: the neighbours still wait for this rank's flag)
  dd_refresh_shifts(h);
  const int w = dd_width(h, false), nown = (int)(h->n - h->nghost);
  if (ev_tick(h, 6)) return -2;
  PeerPlan plan{};
  const double *src = D.recvbuf.p;
  if (D.peer.enabled) {
    PeerCtx &P = D.peer;
    P.seq_f++;
    dd_peer_plan(h, false, (int)(P.seq_f & 1), plan);
    src = peer_inbox(P, P.region, P.capF, P.capR, false, (int)(P.seq_f & 1), P.wmax);
  }
  if (D.nsend > 0) {
    dd_pack_kernel<<<cdiv(D.nsend, 256), 256, 0, h->stream>>>(view(h), h->d_tag.p, D.G, D.nsend, D.send_idx.p, D.send_slot.p, 0, D.ghost_vel, D.sendbuf.p, plan);
    h->kernel_launches++;
  }
  int rc2;
  if (D.peer.enabled) { if ((rc2 = dd_peer_sync(h, false))) return rc2; }
  else if ((rc2 = dd_sendrecv(h, D.nbr_rank, D.sendbuf.p, D.send_cnt, D.recvbuf.p, D.recv_cnt, w))) return rc2;
  if (h->nghost > 0) {
    dd_unpack_kernel<<<cdiv(h->nghost, 256), 256, 0, h->stream>>>(view_all(h), h->d_tag.p, nown, (int)h->nghost, 0, D.ghost_vel, src);
    h->kernel_launches++;
  }
  ev_tock(h);
  return 0;
}

// reverse_comm (newton on): f / torque gathered on the ghosts return to their owners and are added in a fixed order
int dd_reverse(sh_ctx *h) {
  DdCtx &D = h->dd;
  if (!D.peer.enabled && D.nsend == 0 && h->nghost == 0) return 0;
  const int nown = (int)(h->n - h->nghost), ng = (int)h->nghost;
  if (ev_tick(h, 6)) return -2;
  try { D.sendbuf.ensure((size_t)6 * std::max(ng, 1)); D.recvbuf.ensure((size_t)6 * std::max(D.nsend, 1)); }
  catch (std::string &e) { return fail(h, e); }
  PeerPlan plan{};
  const double *src = D.recvbuf.p;
  if (D.peer.enabled) {
    PeerCtx &P = D.peer;
    P.seq_r++;
    dd_peer_plan(h, true, (int)(P.seq_r & 1), plan);
    src = peer_inbox(P, P.region, P.capF, P.capR, true, (int)(P.seq_r & 1), P.wmax);
  }
  if (ng > 0) {
    dd_pack_ghost_forces_kernel<<<cdiv(ng, 256), 256, 0, h->stream>>>(view_all(h), nown, ng, D.sendbuf.p, plan);
    h->kernel_launches++;
  }
  int rc2;
  if (D.peer.enabled) { if ((rc2 = dd_peer_sync(h, true))) return rc2; }
  else if ((rc2 = dd_sendrecv(h, D.nbr_rank, D.sendbuf.p, D.recv_cnt, D.recvbuf.p, D.send_cnt, 6))) return rc2;
  if (D.nsend > 0 && nown > 0) {
    dd_add_returned_forces_kernel<<<cdiv(nown, 256), 256, 0, h->stream>>>(view(h), D.rev_off.p, D.rev_k.p, src);
    h->kernel_launches++;
  }
  ev_tock(h);
  return 0;
}

int dd_rebuild(sh_ctx *h) {
  int rc2;
  if (ev_tick(h, 6)) return -2;
  if ((rc2 = dd_migrate(h))) return rc2;
  if ((rc2 = dd_borders(h))) return rc2;
  ev_tock(h);
  return 0;
}

}  // namespace
