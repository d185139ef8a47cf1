"""Spatial domain decomposition of the SPHERHARM path across ranks (row a12).

One process per GPU (torchrun).  The global box is cut into a px x py x pz brick grid; every rank owns the
atoms inside its brick and keeps GHOST copies of the atoms owned by neighbouring bricks (or periodic
images) that lie within `rc = 2 Rmax + skin` of its faces, edges and corners.  "Newton off": a pair that
straddles a boundary is evaluated on both ranks and each accumulates only into its owned atom, so no force
return trip exists (LAMMPS `newton off`, the setting granular SPHERHARM inputs use).

Per step (device path): sh_step_begin (integrate) -> MAX all-reduce of the rebuild flag -> forward
exchange of ghost x / quat: one CUDA pack kernel, ONE all_to_all_single (NCCL grouped send/recv over
NVLink; 56 B per ghost), one CUDA unpack kernel -> sh_step_end (ghost poses, pair kernel, gather, integrate).
On neighbor-rebuild steps only (host path): owned state comes back to the host, atoms that left the brick
migrate (all_to_all of variable-size records), the border lists are rebuilt and the ghosts re-created.

The class is transport-agnostic: torch.distributed with NCCL (CUDA tensors) on the GPU box, gloo (CPU
tensors) in the CPU tests, where the engine is any object with the sh_* call surface.
"""
import itertools

import numpy as np
import torch
import torch.distributed as dist

OFFSETS = [o for o in itertools.product((-1, 0, 1), repeat=3) if o != (0, 0, 0)]
NREC = 15  # migration record: tag, shape, x3, v3, quat4, angmom3


def proc_grid(nranks, boxlen):
    """Factor nranks into (px,py,pz) minimising the total cut surface for the given box lengths."""
    best, best_cost = (nranks, 1, 1), None
    for px in range(1, nranks + 1):
        if nranks % px:
            continue
        for py in range(1, nranks // px + 1):
            if (nranks // px) % py:
                continue
            pz = nranks // px // py
            sx, sy, sz = boxlen[0] / px, boxlen[1] / py, boxlen[2] / pz
            cost = sx * sy * (pz > 1) + sy * sz * (px > 1) + sx * sz * (py > 1)     # ghost surface per rank
            cost = cost * (1.0 + 1e-3 * (sx * sx + sy * sy + sz * sz) / (sx * sy + sy * sz + sx * sz))  # prefer cubic bricks
            cost = cost + 1e-9 * (px * 100 + py * 10 + pz)     # deterministic tie-break
            if best_cost is None or cost < best_cost:
                best, best_cost = (px, py, pz), cost
    return best


class DomainDecomposition:
    def __init__(self, engine, cfg, rank=None, world=None, comm_device="cpu", rmax=None, com_max=0.0):
        """engine: object with the sh_* call surface (already created on this rank's device).
        cfg: GLOBAL workload dict (identical on every rank); needs cfg['box'] = (lo, hi, periodic)."""
        self.e = engine
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.dev = torch.device(comm_device)
        lo, hi, per = cfg["box"]
        self.lo, self.hi = np.asarray(lo, float), np.asarray(hi, float)
        self.L = self.hi - self.lo
        self.gper = np.asarray(per, int)
        self.pgrid = np.array(proc_grid(self.world, self.L))
        self.g = np.array(np.unravel_index(self.rank, self.pgrid))        # my brick coordinates
        self.sub = self.L / self.pgrid
        self.mylo = self.lo + self.g * self.sub
        self.myhi = self.mylo + self.sub
        self.local_per = [int(self.gper[d] and self.pgrid[d] == 1) for d in range(3)]
        self.cfg = cfg
        self.skin = float(cfg["skin"])
        self._setup_engine(cfg)
        if rmax is None:
            rmax = max(self.e.shape_props(s)["rmax"] for s in range(len(cfg["shapes"])))
            com_max = max(float(np.linalg.norm(self.e.shape_props(s)["com"])) for s in range(len(cfg["shapes"])))
        self.rc = 2.0 * rmax + self.skin + 2.0 * com_max
        for d in range(3):
            if self.pgrid[d] > 1 and self.sub[d] < self.rc:
                raise ValueError("sub-domain thinner than the ghost cutoff in dim %d" % d)
        # initial ownership straight from the global arrays
        x = self._wrap(np.asarray(cfg["x"], float))
        mine = self._owner(x) == self.rank
        n = len(x)
        tag = np.arange(1, n + 1, dtype=np.int64)
        self.tag = tag[mine]
        self.shape = np.asarray(cfg["shape_id"], np.int32)[mine]
        st = dict(x=x[mine], v=np.asarray(cfg["v"], float)[mine], quat=np.asarray(cfg["quat"], float)[mine],
                  angmom=np.asarray(cfg["angmom"], float)[mine])
        self._borders_and_upload(st)

    # ------------------------------------------------------------------ engine configuration
    def _setup_engine(self, cfg):
        e = self.e
        e.set_box(self.lo, self.hi, self.local_per)
        e.set_quadrature(*cfg["grid"])
        ids = [e.add_shape(cfg["lmax"], a, b, cfg["density"]) for (a, b) in cfg["shapes"]]
        k, ex = cfg["coeff"]
        for i in ids:
            for j in ids:
                if j >= i:
                    e.pair_coeff(i, j, k, ex)
        for (pt, nrm, kw, ew) in cfg["walls"]:
            e.add_wall(pt, nrm, kw, ew)
        e.set_gravity(cfg["gravity"])
        e.set_neighbor(cfg["skin"], 1, 1)
        e.set_timestep(cfg["dt"])

    # ------------------------------------------------------------------ geometry helpers
    def _wrap(self, x):
        x = np.array(x, dtype=float, copy=True)
        for d in range(3):
            if self.gper[d]:
                x[:, d] -= self.L[d] * np.floor((x[:, d] - self.lo[d]) / self.L[d])
        return x

    def _owner(self, x):
        gi = np.empty((len(x), 3), dtype=np.int64)
        for d in range(3):
            c = np.floor((x[:, d] - self.lo[d]) / self.sub[d]).astype(np.int64)
            gi[:, d] = np.clip(c, 0, self.pgrid[d] - 1)
        return np.ravel_multi_index((gi[:, 0], gi[:, 1], gi[:, 2]), self.pgrid)

    def _neighbor(self, o):
        """(rank, shift) of the brick at offset o, or None (non-periodic edge / undecomposed dim)."""
        gg, shift = self.g.copy(), np.zeros(3)
        for d in range(3):
            if o[d] == 0:
                continue
            if self.pgrid[d] == 1:
                return None
            gg[d] += o[d]
            if gg[d] >= self.pgrid[d]:
                if not self.gper[d]:
                    return None
                gg[d] -= self.pgrid[d]; shift[d] = -self.L[d]
            elif gg[d] < 0:
                if not self.gper[d]:
                    return None
                gg[d] += self.pgrid[d]; shift[d] = self.L[d]
        return int(np.ravel_multi_index(tuple(gg), self.pgrid)), shift

    # ------------------------------------------------------------------ collectives (variable size)
    def _alltoall_rows(self, rows_by_dest, ncol, dtype=torch.float64):
        """rows_by_dest: list[world] of (k_r x ncol) numpy arrays -> list[world] of received arrays."""
        counts = torch.tensor([len(r) for r in rows_by_dest], dtype=torch.int64, device=self.dev)
        rcounts = torch.empty_like(counts)
        dist.all_to_all_single(rcounts, counts)
        sc, rc = counts.tolist(), rcounts.tolist()
        send = np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, ncol) for r in rows_by_dest], axis=0) \
            if sum(sc) else np.zeros((0, ncol))
        tsend = torch.from_numpy(np.ascontiguousarray(send)).to(self.dev)
        trecv = torch.empty((sum(rc), ncol), dtype=torch.float64, device=self.dev)
        dist.all_to_all_single(trecv, tsend, output_split_sizes=rc, input_split_sizes=sc)
        out, o = [], 0
        got = trecv.cpu().numpy()
        for r in range(self.world):
            out.append(got[o:o + rc[r]]); o += rc[r]
        return out, sc, rc

    # ------------------------------------------------------------------ borders (ghost construction)
    def _borders_and_upload(self, st):
        """st: owned state dict (x wrapped into the global box).  Builds send lists, exchanges ghosts,
        uploads owned+ghost atoms to the engine."""
        x = st["x"]
        nloc = len(x)
        send_rows = [[] for _ in range(self.world)]
        send_idx = [[] for _ in range(self.world)]
        send_shift = [[] for _ in range(self.world)]
        # only atoms in the shell within rc of a decomposed face can be ghosts of anybody
        near_hi = [x[:, d] >= self.myhi[d] - self.rc if self.pgrid[d] > 1 else None for d in range(3)]
        near_lo = [x[:, d] < self.mylo[d] + self.rc if self.pgrid[d] > 1 else None for d in range(3)]
        shell_mask = np.zeros(nloc, dtype=bool)
        for d in range(3):
            if self.pgrid[d] > 1:
                shell_mask |= near_hi[d] | near_lo[d]
        shell = np.nonzero(shell_mask)[0]
        sh_hi = [near_hi[d][shell] if near_hi[d] is not None else None for d in range(3)]
        sh_lo = [near_lo[d][shell] if near_lo[d] is not None else None for d in range(3)]
        for o in OFFSETS:
            nb = self._neighbor(o)
            if nb is None:
                continue
            dest, shift = nb
            m = np.ones(len(shell), dtype=bool)
            for d in range(3):
                if o[d] == 1:
                    m &= sh_hi[d]
                elif o[d] == -1:
                    m &= sh_lo[d]
            idx = shell[m]
            if len(idx) == 0:
                continue
            send_idx[dest].append(idx)
            send_shift[dest].append(np.tile(shift, (len(idx), 1)))
            rows = np.empty((len(idx), 9))
            rows[:, 0] = self.tag[idx]; rows[:, 1] = self.shape[idx]
            rows[:, 2:5] = x[idx] + shift; rows[:, 5:9] = st["quat"][idx]
            send_rows[dest].append(rows)
        rows_by_dest = [np.concatenate(r, axis=0) if r else np.zeros((0, 9)) for r in send_rows]
        got, sc, rc = self._alltoall_rows(rows_by_dest, 9)
        ghosts = np.concatenate(got, axis=0) if sum(rc) else np.zeros((0, 9))
        self.send_counts, self.recv_counts = sc, rc
        self.nlocal, self.nghost = nloc, len(ghosts)
        idx_all = np.concatenate([np.concatenate(i) for i in send_idx if i]) if sum(sc) else np.zeros(0, np.int64)
        shift_all = np.concatenate([np.concatenate(s) for s in send_shift if s], axis=0) if sum(sc) else np.zeros((0, 3))
        self.nsend = int(sum(sc))
        # device-side send lists and buffers for the per-step forward exchange
        self.t_idx = torch.from_numpy(idx_all.astype(np.int32)).to(self.dev)
        self.t_shift = torch.from_numpy(np.ascontiguousarray(shift_all)).to(self.dev)
        self.t_send = torch.empty((max(1, self.nsend), 7), dtype=torch.float64, device=self.dev)
        self.t_recv = torch.empty((max(1, self.nghost), 7), dtype=torch.float64, device=self.dev)
        self.ghost_tag = ghosts[:, 0].astype(np.int64)
        allx = np.concatenate([x, ghosts[:, 2:5]], axis=0)
        allq = np.concatenate([st["quat"], ghosts[:, 5:9]], axis=0)
        allv = np.concatenate([st["v"], np.zeros((self.nghost, 3))], axis=0)
        allL = np.concatenate([st["angmom"], np.zeros((self.nghost, 3))], axis=0)
        allshape = np.concatenate([self.shape, ghosts[:, 1].astype(np.int32)])
        alltag = np.concatenate([self.tag, self.ghost_tag])
        self.e.set_atoms(allshape, allx, allv, allq, allL, tag=alltag)
        if hasattr(self.e, "set_ghost_count"):
            self.e.set_ghost_count(self.nghost)

    # ------------------------------------------------------------------ per-step forward exchange (device)
    def forward(self):
        if self.nsend:
            self.e.pack_atoms(self.nsend, self.t_idx.data_ptr(), self.t_shift.data_ptr(), self.t_send.data_ptr())
        dist.all_to_all_single(self.t_recv[:self.nghost], self.t_send[:self.nsend],
                               output_split_sizes=self.recv_counts, input_split_sizes=self.send_counts)
        if self.dev.type == "cuda":
            torch.cuda.current_stream().synchronize()
        if self.nghost:
            self.e.unpack_ghosts(self.nlocal, self.nghost, self.t_recv.data_ptr())

    # ------------------------------------------------------------------ rebuild: migrate + borders (host)
    def rebuild(self):
        st = self.e.get_atoms(("x", "v", "quat", "angmom"))
        nl = self.nlocal
        x = self._wrap(st["x"][:nl])
        owner = self._owner(x)
        rec = np.empty((nl, NREC))
        rec[:, 0] = self.tag; rec[:, 1] = self.shape; rec[:, 2:5] = x; rec[:, 5:8] = st["v"][:nl]
        rec[:, 8:12] = st["quat"][:nl]; rec[:, 12:15] = st["angmom"][:nl]
        # atoms that stay never touch the transport: only the migrants are exchanged
        stay = owner == self.rank
        rows_by_dest = [rec[owner == r] if r != self.rank else rec[:0] for r in range(self.world)]
        got, _, _ = self._alltoall_rows(rows_by_dest, NREC)
        mine = np.concatenate([rec[stay]] + got, axis=0)
        order = np.argsort(mine[:, 0], kind="stable")          # deterministic local order: by tag
        mine = mine[order]
        self.tag = mine[:, 0].astype(np.int64)
        self.shape = mine[:, 1].astype(np.int32)
        self._borders_and_upload(dict(x=mine[:, 2:5], v=mine[:, 5:8], quat=mine[:, 8:12], angmom=mine[:, 12:15]))

    # ------------------------------------------------------------------ time stepping
    def setup(self):
        self.e.compute_forces()

    def step(self):
        flag = self.e.step_begin()
        t = torch.tensor([flag], dtype=torch.int32, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        flag = int(t.item())
        if flag:
            self.rebuild()
        else:
            self.forward()
        self.e.step_end(flag)
        return flag

    def run(self, nsteps):
        nreb = 0
        for _ in range(nsteps):
            nreb += self.step()
        self.e.synchronize()
        return nreb

    # ------------------------------------------------------------------ gather to rank 0 (tests / output)
    def gather_owned(self, fields=("x", "v", "quat", "angmom", "f", "torque")):
        st = self.e.get_atoms()
        ncol = {"x": 3, "v": 3, "quat": 4, "angmom": 3, "f": 3, "torque": 3}
        rows = np.concatenate([self.tag[:, None].astype(float)] + [st[k][:self.nlocal] for k in fields], axis=1)
        w = rows.shape[1]
        rows_by_dest = [rows if r == 0 else np.zeros((0, w)) for r in range(self.world)]
        got, _, _ = self._alltoall_rows(rows_by_dest, w)
        if self.rank != 0:
            return None
        allr = np.concatenate(got, axis=0)
        allr = allr[np.argsort(allr[:, 0], kind="stable")]
        out, o = {"tag": allr[:, 0].astype(np.int64)}, 1
        for k in fields:
            out[k] = allr[:, o:o + ncol[k]]; o += ncol[k]
        return out


# ---------------------------------------------------------------------- in-library decomposition (NCCL inside libshgpu)
def native_engine(pkg, cfg, device, tuning=None, pgrid=None, shear=0.0):
    """One engine per rank whose decomposition (migration, borders, ghost exchange) runs inside libshgpu (sh_dd_*).
    torch.distributed only carries the 128-byte NCCL id from rank 0 to the others."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    uid = [pkg.ShGpu.dd_unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    sim = pkg.ShGpu(device=device)
    for k, v in (tuning or {}).items():
        sim.set_tuning(k, v)
    sim.dd_init(rank, world, uid[0], pgrid)
    if shear:
        sim.set_shear(shear)
    pkg.workloads.apply(sim, cfg)
    return sim


def gather_owned_native(sim, fields=("x", "v", "quat", "angmom", "f", "torque")):
    """Owned atoms of every rank, sorted by tag, on rank 0 (tests / output)."""
    info = sim.dd_info()
    nl = info["nlocal"]
    st = sim.get_atoms(fields)
    mine = {k: st[k][:nl] for k in fields}
    mine["tag"] = sim.get_tags()[:nl]
    if dist.is_initialized() and dist.get_world_size() > 1:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, mine)
        if dist.get_rank() != 0:
            return None
    else:
        parts = [mine]
    tag = np.concatenate([p["tag"] for p in parts])
    order = np.argsort(tag, kind="stable")
    out = {"tag": tag[order]}
    for k in fields:
        out[k] = np.concatenate([p[k] for p in parts], axis=0)[order]
    return out
