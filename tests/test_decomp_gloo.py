"""World-size-2/4 gloo tests (CPU) of the domain-decomposition host logic: ownership, ghost construction
(faces/edges/corners, periodic shifts), migration.  The local engine is the CPU oracle; the decomposed
per-atom forces must equal the single-domain forces."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
ROOT = sys.argv[1]; periodic = int(sys.argv[2]); out = sys.argv[3]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import shpkg, oracle_py as O
pkg = shpkg.load(); W = pkg.workloads; D = pkg.load_decomp()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cells = (5, 5, 5) if world == 8 else (4, 3, 3)      # 2x2x2 bricks: in a periodic box the +1 and -1 neighbours coincide
cfg = W.packing(cells, 6, (12, 24), nshapes=3, seed=9, periodic=bool(periodic), name="dd", skin=0.1)
ref = None
if rank == 0:
    o = O.Oracle(threads=2); W.apply(o, cfg); o.compute_forces(); ref = o.get_atoms()
dd = D.DomainDecomposition(O.Oracle(threads=1 if world == 8 else 2), cfg, comm_device="cpu")
if world == 8:
    assert tuple(dd.pgrid) == (2, 2, 2), dd.pgrid
dd.setup()
got = dd.gather_owned(("x", "f", "torque"))
counts = [None] * world
dist.all_gather_object(counts, (dd.nlocal, dd.nghost))
# move the atoms (as a few timesteps would) and re-decompose: exercises migration
st = dd.e.get_atoms(("x", "v", "quat", "angmom"))
rng = np.random.default_rng(100 + rank)
xn = st["x"].copy(); xn[:dd.nlocal] += rng.normal(0, 0.4, size=(dd.nlocal, 3))
dd.e.set_atoms(np.concatenate([dd.shape, np.zeros(dd.nghost, np.int32)]), xn, st["v"], st["quat"], st["angmom"])
moved = dd.gather_owned(("x",))     # positions by tag before migration (owned rows only)
dd.rebuild(); dd.setup()
got2 = dd.gather_owned(("x", "f", "torque"))
counts2 = [None] * world
dist.all_gather_object(counts2, (dd.nlocal, dd.nghost))
if rank == 0:
    n = len(cfg["x"])
    assert sum(c[0] for c in counts) == n and sum(c[0] for c in counts2) == n, (counts, counts2)
    assert all(c[1] > 0 for c in counts), counts
    fs = np.abs(ref["f"]).max()
    assert np.abs(got["f"] - ref["f"]).max() <= 1e-11 * fs, np.abs(got["f"] - ref["f"]).max()
    assert np.abs(got["torque"] - ref["torque"]).max() <= 1e-11 * fs
    # after the move: reference = single-domain oracle on the moved positions
    o = O.Oracle(threads=2); cfg2 = dict(cfg); cfg2["x"] = moved["x"]; W.apply(o, cfg2); o.compute_forces(); ref2 = o.get_atoms()
    L = np.asarray(cfg["box"][1]) - np.asarray(cfg["box"][0])
    dx = got2["x"] - moved["x"]
    if periodic:
        dx -= L * np.rint(dx / L)
    assert np.abs(dx).max() < 1e-12
    fs2 = np.abs(ref2["f"]).max()
    assert np.abs(got2["f"] - ref2["f"]).max() <= 1e-10 * fs2, np.abs(got2["f"] - ref2["f"]).max()
    open(out, "w").write("ok %s %s" % (counts, counts2))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world,periodic", [(2, 1), (2, 0), (4, 1), (8, 1), (8, 0)])
def test_decomposed_forces_equal_single_domain(tmp_path, world, periodic):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    out = tmp_path / "ok.txt"
    env = dict(os.environ, OMP_NUM_THREADS="1" if world == 8 else "2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world * 2 + periodic), str(w), ROOT, str(periodic), str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert out.read_text().startswith("ok")


WORKER_TRAJ = r'''
import os, sys
import numpy as np
import torch.distributed as dist
ROOT = sys.argv[1]; out = sys.argv[2]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import shpkg, oracle_py as O
pkg = shpkg.load(); W = pkg.workloads; D = pkg.load_decomp()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cfg = W.packing((4, 3, 3), 6, (12, 24), nshapes=3, seed=9, periodic=True, name="ddtraj", skin=0.04, vel_sigma=0.6, dt=5e-4, nn_frac=1.8)
nsteps = 160
dd = D.DomainDecomposition(O.Oracle(threads=2), cfg, comm_device="cpu")
dd.setup()
nreb = dd.run(nsteps)                    # step_begin -> flag all-reduce -> forward exchange / migrate+borders -> step_end
got = dd.gather_owned(("x", "v", "quat", "angmom"))
if rank == 0:
    o = O.Oracle(threads=2); W.apply(o, cfg); o.run(nsteps); ref = o.get_atoms()
    L = np.asarray(cfg["box"][1])
    dx = got["x"] - ref["x"]; dx -= L * np.rint(dx / L)
    assert np.abs(dx).max() <= 1e-10, np.abs(dx).max()
    for k, tol in (("v", 1e-9), ("quat", 1e-10), ("angmom", 1e-9)):
        assert np.abs(got[k] - ref[k]).max() <= tol * max(1.0, np.abs(ref[k]).max()), (k, np.abs(got[k] - ref[k]).max())
    assert np.abs(ref["angmom"]).max() > 1e-3, "no collisions happened"
    assert nreb >= 3, "the run must exercise migration / ghost re-creation (got %d rebuilds)" % nreb
    open(out, "w").write("ok rebuilds=%d" % nreb)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4])
def test_decomposed_trajectory_equals_single_domain(tmp_path, world):
    """The full multi-rank step loop of decomp.py on CPU (gloo): ghost exchange every step, migration and ghost
    re-creation on rebuild steps, against a single-domain oracle run."""
    w = tmp_path / "worker_traj.py"
    w.write_text(WORKER_TRAJ)
    out = tmp_path / "ok.txt"
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29680 + world), str(w), ROOT, str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert out.read_text().startswith("ok")
