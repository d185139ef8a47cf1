// dd_types.h — state of the in-library domain decomposition (see dd_host.cuh)
#pragma once
#include <nccl.h>

#include <string>
#include <vector>

#include "decomp_kernels.cuh"

namespace shgpu {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};


// Per-step ghost exchange over NVLink peer memory: every rank owns one device region (cudaMalloc, exported with
// cudaIpcGetMemHandle; same-process ranks use the pointer directly) holding two inboxes, each double-buffered by the parity
// of the exchange, and two flag rows indexed by the writer's rank:
//   F[2][capF * wmax]  ghost records pushed by the neighbours' pack kernel (forward)
//   R[2][capR * 6]     force / torque returned for the atoms this rank sent (reverse, newton on)
//   flags[2][DD_MAX_RANKS]  sequence number of the last completed push per writer
// A neighbour's pack kernel stores straight into the inbox (NVLink stores); a one-warp kernel then publishes the sequence
// number to every neighbour and waits for theirs.  NCCL stays in charge of everything that happens on rebuild steps.
struct PeerCtx {
  bool enabled = false, want = true;
  double *region = nullptr;
  long long capF = 0, capR = 0;                 // records
  int wmax = 13;
  std::vector<double *> base;                    // per neighbour (index into nbr_rank): that rank's region as seen from here
  std::vector<long long> pcapF, pcapR;           // its capacities
  std::vector<void *> opened;                    // cudaIpcOpenMemHandle results to close
  DevBuf<long long> d_off;                       // [0..31] my record offset in neighbour k's F inbox, [32..63] in its R inbox
  unsigned long long seq_f = 0, seq_r = 0;
  int *d_err = nullptr;                          // device flag: a wait timed out
};

struct DdCtx {
  PeerCtx peer;
  bool on = false, geometry_ok = false, borders_ok = false, pgrid_set = false;
  int rank = 0, nranks = 1;
  NcclApi *nccl = nullptr;
  ncclComm_t comm = nullptr;
  int pgrid[3] = {1, 1, 1}, g[3] = {0, 0, 0};
  double glo[3] = {0, 0, 0}, ghi[3] = {0, 0, 0}, glen[3] = {0, 0, 0};
  int gper[3] = {0, 0, 0};
  DdGeom G{};
  std::vector<int> nbr_rank, slot_lo, slot_hi;   // distinct neighbour ranks (ascending) and their slot ranges
  std::vector<int> mig_rank;                     // the same without this rank (migration targets)
  std::vector<int> send_cnt, recv_cnt;           // ghosts per neighbour rank of the forward exchange
  std::vector<double> base_shift;                // periodic shift per slot (without the Lees-Edwards offset)
  int nsend = 0, ghost_vel = 0;
  bool newton = true;                            // cross-rank pairs evaluated once, reactions returned (reverse comm)
  DevBuf<int> rev_cnt, rev_off, rev_k;           // per owned atom: its positions in the send list (ascending slot order)
  bool self_ghosts = false;                      // test knob: periodic dims get ghost images even when undivided
  double le_rate = 0, le_off_build = 0, le_time_build = 0;   // Lees-Edwards: shear rate, image offset at the last rebuild
  DevBuf<int> flag, pos, order, d_int, send_idx, send_slot, shape2;
  DevBuf<double> sendbuf, recvbuf, x2, v2, q2, L2;
  DevBuf<long long> tag2;
  int stride2 = 0, last_nstay = 0;
  int *h_int = nullptr;                          // pinned, 256 ints
  long long h_off[64] = {0};                     // my offsets in the neighbours' inboxes (forward 0..31, reverse 32..63)
  int64_t migrated_out = 0, migrated_in = 0, border_builds = 0;
};

}  // namespace shgpu
