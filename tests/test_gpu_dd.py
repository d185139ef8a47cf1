"""GPU tests of the in-library domain decomposition (sh_dd_*, csrc/dd_host.cuh + decomp_kernels.cuh).

One GPU: with "dd_self_ghosts" the periodic dimensions are served by ghost images of the rank's own atoms, so migration
(wrap + compaction), border lists, the per-step ghost exchange and the lagged neighbor decision all run — against the
plain periodic engine, which must give the same forces (1e-11) and the same trajectory (1e-9 over 300 steps with rebuilds).
Two GPUs (`gpurun --gpus 2`): NCCL inside libshgpu, torchrun only carries the 128-byte id.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import shpkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = shpkg.load()
W = pkg.workloads
pytestmark = pytest.mark.gpu


def _cfg():
    cfg = W.packing((6, 4, 4), 20, (32, 64), nshapes=4, seed=21, periodic=True, name="dd", skin=0.03, vel_sigma=0.5, dt=4e-4)
    cfg["v"] = cfg["v"] + np.array([8.0, 1.0, 0.0])      # a drift, so that atoms cross brick / box boundaries (migration)
    return cfg


def _sorted_owned(sim):
    info = sim.dd_info()
    nl = info["nlocal"]
    st = sim.get_atoms()
    tag = sim.get_tags()[:nl]
    o = np.argsort(tag)
    return {k: v[:nl][o] for k, v in st.items()}, info


@pytest.mark.parametrize("variant,sync,newton", [(0, 0, 1), (16, 0, 1), (16, 1, 0), (0, 1, 0), (16, 0, 0)])
def test_self_ghost_decomposition_equals_periodic_engine(variant, sync, newton):
    """newton 1 (default): a pair with a ghost is evaluated once and the reaction returns to the owner (reverse
    communication); newton 0: both sides evaluate it and keep their own half."""
    cfg = _cfg()
    ref = pkg.ShGpu(); W.apply(ref, cfg); ref.set_pair_tuning(0, 0, variant)
    ref.set_tuning("sync_rebuild", 1)
    dd = pkg.ShGpu(); dd.set_tuning("dd_self_ghosts", 1); dd.set_tuning("sync_rebuild", sync); dd.set_tuning("newton", newton); dd.dd_init(0, 1)
    W.apply(dd, cfg); dd.set_pair_tuning(0, 0, variant)
    ref.compute_forces(); dd.compute_forces()
    r0 = ref.get_atoms(); d0, info = _sorted_owned(dd)
    assert info["nlocal"] == len(cfg["x"]) and info["nghost"] > 0
    cr, cd = ref.get_counters()["pair_evals"], dd.get_counters()["pair_evals"]
    assert (cd == cr) if newton else (cd > cr), (cr, cd)      # newton on: no pair is evaluated twice
    fs = np.abs(r0["f"]).max()
    assert np.abs(d0["f"] - r0["f"]).max() <= 1e-11 * fs
    assert np.abs(d0["torque"] - r0["torque"]).max() <= 1e-11 * fs
    nsteps = 300
    ref.run(nsteps); dd.run(nsteps)
    r1 = ref.get_atoms(); d1, info = _sorted_owned(dd)
    assert info["border_builds"] >= 3, info          # rebuilds (migration + borders) happened
    if variant == 16:                                # the candidate cache survived them (carried over by tag)
        cs = dd.get_cache_stats()
        assert cs["cache_remaps"] >= 2, cs
    L = np.asarray(cfg["box"][1]) - np.asarray(cfg["box"][0])
    dx = d1["x"] - r1["x"]; dx -= L * np.rint(dx / L)
    assert np.abs(dx).max() <= 1e-9, np.abs(dx).max()
    for k, tol in (("v", 1e-8), ("quat", 1e-9), ("angmom", 1e-8), ("f", 1e-7)):
        assert np.abs(d1[k] - r1[k]).max() <= tol * max(1.0, np.abs(r1[k]).max()), k
    ref.close(); dd.close()


def test_lagged_neighbor_decision_matches_classic():
    """sh_run with the one-step-ahead rebuild prediction (default) against the classic per-step flag read-back."""
    cfg = _cfg()
    out = []
    for sync in (1, 0):
        g = pkg.ShGpu(); W.apply(g, cfg); g.set_tuning("sync_rebuild", sync); g.compute_forces(); g.run(400)
        out.append((g.get_atoms(), g.get_counters()["neighbor_builds"])); g.close()
    (a, na), (b, nb) = out
    assert na >= 3 and nb >= 3
    # the prediction rebuilds up to two steps earlier each time (with this fast drift a rebuild comes every ~5 steps)
    assert na - 2 <= nb <= 2 * na + 2, (na, nb)
    L = np.asarray(cfg["box"][1]) - np.asarray(cfg["box"][0])
    dx = a["x"] - b["x"]; dx -= L * np.rint(dx / L)
    assert np.abs(dx).max() <= 1e-9


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = sys.argv[1]; out = sys.argv[2]
sys.path.insert(0, ROOT)
import shpkg
pkg = shpkg.load(); W = pkg.workloads; D = pkg.load_decomp()
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cfg = W.packing((6, 4, 4), 20, (32, 64), nshapes=4, seed=21, periodic=True, name="mr", skin=0.03, vel_sigma=0.5, dt=4e-4)
cfg["v"] = cfg["v"] + np.array([8.0, 1.0, 0.0])      # a drift, so that atoms cross the brick boundaries (migration)
nsteps = 300
for variant in (16, 0):
    sim = D.native_engine(pkg, cfg, local)
    sim.set_pair_tuning(0, 0, variant)
    sim.compute_forces()
    f0 = D.gather_owned_native(sim, ("f", "torque"))
    sim.run(nsteps)
    got = D.gather_owned_native(sim, ("x", "v", "quat", "angmom", "f"))
    info = sim.dd_info()
    if rank == 0:
        g = pkg.ShGpu(device=local); W.apply(g, cfg); g.set_pair_tuning(0, 0, variant); g.compute_forces(); r0 = g.get_atoms()
        fs = np.abs(r0["f"]).max()
        assert np.abs(f0["f"] - r0["f"]).max() <= 1e-11 * fs, np.abs(f0["f"] - r0["f"]).max() / fs
        assert np.abs(f0["torque"] - r0["torque"]).max() <= 1e-11 * fs
        g.run(nsteps); r1 = g.get_atoms()
        L = np.asarray(cfg["box"][1]) - np.asarray(cfg["box"][0])
        dx = got["x"] - r1["x"]; dx -= L * np.rint(dx / L)
        assert np.abs(dx).max() <= 1e-9, np.abs(dx).max()
        for k, tol in (("v", 1e-8), ("quat", 1e-9), ("angmom", 1e-8)):
            assert np.abs(got[k] - r1[k]).max() <= tol * max(1.0, np.abs(r1[k]).max()), k
        assert info["border_builds"] >= 3 and info["migrated"] >= 1, info
        g.close()
    sim.close()
if rank == 0:
    open(out, "w").write("ok %s" % (info,))
dist.destroy_process_group()
'''


def test_two_gpu_native_decomposition_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    out = tmp_path / "ok.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29651", str(w), ROOT, str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    print(out.read_text())
    assert out.read_text().startswith("ok")


@pytest.mark.parametrize("variant", [0, 16])
def test_lees_edwards_shear_box_vs_oracle(variant):
    """BASELINE configs[3] physics at test size: periodic box sheared by Lees-Edwards images (sh_set_shear) against the
    oracle's sheared minimum image, forces at several times and the trajectory in between."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    rate = 0.6
    cfg = W.shear_box(W.packing((5, 4, 4), 20, (32, 64), nshapes=4, seed=33, periodic=True, name="le", skin=0.04, vel_sigma=0.3, dt=4e-4), rate)
    cfg["v"] = cfg["v"] + np.array([0.0, 8.0, 0.0])      # a drift along the gradient direction: atoms cross the sheared boundary
    lo, hi, _ = cfg["box"]
    L = np.asarray(hi) - np.asarray(lo)
    g = pkg.ShGpu(); W.apply(g, cfg); g.set_pair_tuning(0, 0, variant)
    o = O.Oracle(threads=8); W.apply(o, cfg)
    t = 0.0
    for nsteps in (0, 120, 180):
        if nsteps:
            g.run(nsteps); o.run(nsteps); t += nsteps * cfg["dt"]
        else:
            g.compute_forces(); o.compute_forces()
        gs, info = _sorted_owned(g)
        os_ = o.get_atoms()
        off, vs = rate * L[1] * t, rate * L[1]
        ny = np.rint((os_["x"][:, 1] - gs["x"][:, 1]) / L[1])           # the oracle never wraps: it holds image +ny
        dx = os_["x"] - gs["x"]
        dx[:, 1] -= ny * L[1]; dx[:, 0] -= ny * off
        dx[:, 0] -= L[0] * np.rint(dx[:, 0] / L[0]); dx[:, 2] -= L[2] * np.rint(dx[:, 2] / L[2])
        assert np.abs(dx).max() <= 1e-8, (nsteps, np.abs(dx).max())
        dv = os_["v"] - gs["v"]; dv[:, 0] -= ny * vs
        assert np.abs(dv).max() <= 1e-7 * max(1.0, np.abs(os_["v"]).max()), (nsteps, np.abs(dv).max())
        fs = max(np.abs(os_["f"]).max(), 1e-30)
        assert fs > 1.0
        assert np.abs(os_["f"] - gs["f"]).max() <= 1e-7 * fs, (nsteps, np.abs(os_["f"] - gs["f"]).max() / fs)
        assert np.abs(os_["torque"] - gs["torque"]).max() <= 1e-7 * fs
    assert info["nghost"] > 0 and info["border_builds"] >= 3
    assert np.abs(ny).max() >= 1            # some atoms did cross the sheared boundary
    g.close(); o.close()


def test_self_ghost_decomposition_with_dissipation():
    """Ghosts carry v and angmom when the contact law depends on velocities (13 doubles per ghost instead of 7)."""
    cfg = _cfg()
    rng = np.random.default_rng(3)
    cfg["angmom"] = rng.normal(0, 0.5, size=cfg["angmom"].shape)
    cfg["dissipation"] = (3.0, 2.0, 0.4)
    ref = pkg.ShGpu(); W.apply(ref, cfg)
    dd = pkg.ShGpu(); dd.set_tuning("dd_self_ghosts", 1); dd.dd_init(0, 1); W.apply(dd, cfg)
    ref.compute_forces(); dd.compute_forces()
    r0 = ref.get_atoms(); d0, info = _sorted_owned(dd)
    fs = np.abs(r0["f"]).max()
    assert np.abs(d0["f"] - r0["f"]).max() <= 1e-11 * fs and np.abs(d0["torque"] - r0["torque"]).max() <= 1e-11 * fs
    ref.run(200); dd.run(200)
    r1 = ref.get_atoms(); d1, info = _sorted_owned(dd)
    L = np.asarray(cfg["box"][1]) - np.asarray(cfg["box"][0])
    dx = d1["x"] - r1["x"]; dx -= L * np.rint(dx / L)
    assert np.abs(dx).max() <= 1e-9 and np.abs(d1["angmom"] - r1["angmom"]).max() <= 1e-8
    sr, sd = ref.get_stress(), dd.get_stress()
    assert np.abs(sr["virial"] - sd["virial"]).max() <= 1e-8 * np.abs(sr["virial"]).max()
    ref.close(); dd.close()
