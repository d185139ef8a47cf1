"""GPU multi-rank test (needs >= 2 GPUs; run with `gpurun --gpus 2`): decomposed NCCL run vs single GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = sys.argv[1]; out = sys.argv[2]
sys.path.insert(0, ROOT)
import shpkg
pkg = shpkg.load(); W = pkg.workloads; D = pkg.load_decomp()
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
cfg = W.packing((6, 4, 4), 20, (32, 64), nshapes=4, seed=21, periodic=True, name="mr", skin=0.03, vel_sigma=0.5, dt=4e-4)
nsteps = 300
eng = pkg.ShGpu(device=local); eng.set_pair_tuning(0, 0, 16)   # force the split pipeline + candidate cache with ghosts
dd = D.DomainDecomposition(eng, cfg, comm_device="cuda")
dd.setup()
f0 = dd.gather_owned(("f", "torque"))
nreb = dd.run(nsteps)
got = dd.gather_owned(("x", "v", "quat", "angmom", "f"))
if rank == 0:
    g = pkg.ShGpu(device=local); W.apply(g, cfg); g.set_pair_tuning(0, 0, 16); g.compute_forces(); r0 = g.get_atoms()
    fs = np.abs(r0["f"]).max()
    assert np.abs(f0["f"] - r0["f"]).max() <= 1e-11 * fs, np.abs(f0["f"] - r0["f"]).max() / fs
    assert np.abs(f0["torque"] - r0["torque"]).max() <= 1e-11 * fs
    g.run(nsteps); r1 = g.get_atoms()
    L = np.asarray(cfg["box"][1])
    dx = got["x"] - r1["x"]; dx -= L * np.rint(dx / L)
    assert np.abs(dx).max() <= 1e-9, np.abs(dx).max()
    for k, tol in (("v", 1e-8), ("quat", 1e-9), ("angmom", 1e-8)):
        assert np.abs(got[k] - r1[k]).max() <= tol * max(1.0, np.abs(r1[k]).max()), k
    assert nreb >= 2, "test must exercise migration/rebuild (got %d rebuilds)" % nreb
    open(out, "w").write("ok rebuilds=%d nlocal=%d nghost=%d" % (nreb, dd.nlocal, dd.nghost))
dist.destroy_process_group()
'''


def test_two_gpu_trajectory_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    out = tmp_path / "ok.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29650", str(w), ROOT, str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    print(out.read_text())
    assert out.read_text().startswith("ok")
