"""GPU test of the shlmp input-script front-end (C++ host over the C-ABI): the example scripts must
reproduce the Python/ctypes path driven with the same parameters."""
import os
import subprocess

import numpy as np
import pytest

import shpkg

pkg = shpkg.load()
W = pkg.workloads
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def read_last_dump(fn):
    lines = open(fn).read().splitlines()
    last = max(i for i, l in enumerate(lines) if l.startswith("ITEM: ATOMS"))
    rows = np.array([[float(v) for v in l.split()] for l in lines[last + 1:] if l and not l.startswith("ITEM")])
    return rows[np.argsort(rows[:, 0])]


def axis_angle(ax, deg):
    ax = np.asarray(ax, float) / np.linalg.norm(ax)
    th = np.deg2rad(deg)
    return np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * ax])


def run_shlmp(script, tmp_path, extra=()):
    from lammps_spherharm_b200 import build as b
    exe = b.build_host()
    ex = os.path.join(ROOT, "examples")
    for f in os.listdir(ex):
        if not f.startswith("dump.") and not os.path.lexists(tmp_path / f):
            os.symlink(os.path.join(ex, f), tmp_path / f)
    res = subprocess.run([exe, "-in", script] + list(extra), cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


def test_two_particle_script_matches_capi(tmp_path):
    out = run_shlmp("in.two_particle", tmp_path)
    assert "Loop time" in out and "Step" in out
    rows = read_last_dump(tmp_path / "dump.two_particle")
    cfg = W.config1_two_particle()
    cfg["x"] = np.array([[-1.0, 0.05, 0.0], [1.0, -0.05, 0.02]])
    cfg["quat"] = np.array([axis_angle([1, 1, 0], 30), axis_angle([0, 1, 1], 75)])
    cfg["dt"] = 5e-4
    g = pkg.ShGpu(); W.apply(g, cfg); g.run(1200)
    at = g.get_atoms()
    assert np.abs(at["angmom"]).max() > 1e-6
    assert np.abs(rows[:, 2:5] - at["x"]).max() < 1e-12
    assert np.abs(rows[:, 5:9] - at["quat"]).max() < 1e-12
    assert np.abs(rows[:, 9:12] - at["v"]).max() < 1e-11


def test_wall_settle_script_runs_and_settles(tmp_path):
    out = run_shlmp("in.wall_settle", tmp_path)
    rows = read_last_dump(tmp_path / "dump.wall_settle")
    assert len(rows) == 1000
    assert rows[:, 4].min() > 0.3          # nobody fell through the wall
    th = [l.split() for l in out.splitlines() if l.strip() and l.split()[0].isdigit()]
    assert len(th) >= 5 and float(th[-1][3]) > 0.0   # contact energy present at the end


def test_script_errors_are_reported(tmp_path):
    from lammps_spherharm_b200 import build as b
    exe = b.build_host()
    bad = tmp_path / "in.bad"
    bad.write_text("atom_style spherharm 20 32 64 missing.sh\npair_style lj/cut 2.5\n")
    res = subprocess.run([exe, "-in", str(bad)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 1 and "ERROR: Unknown pair style lj/cut" in res.stderr


def test_shear_box_script_matches_capi(tmp_path):
    """fix deform ... xy erate (Lees-Edwards) through the script front-end equals the ctypes path (sh_set_shear)."""
    out = run_shlmp("in.shear_box", tmp_path)
    assert "Loop time" in out
    rows = read_last_dump(tmp_path / "dump.shear_box")
    cfg = W.shear_box(W.packing((5, 4, 4), 20, (32, 64), nshapes=1, seed=33, periodic=True, vel_sigma=0.3), 0.6)
    cfg["skin"], cfg["dt"] = 0.04, 4e-4
    g = pkg.ShGpu(); W.apply(g, cfg); g.run(300)
    at = g.get_atoms(); tag = g.get_tags()
    nl = g.dd_info()["nlocal"]
    o = np.argsort(tag[:nl])
    assert len(rows) == nl == 320
    assert np.abs(rows[:, 2:5] - at["x"][:nl][o]).max() < 1e-9
    assert np.abs(rows[:, 9:12] - at["v"][:nl][o]).max() < 1e-8


def test_shlmp_two_gpus_reproduces_one_gpu(tmp_path):
    """shlmp -gpus 2 (two rank threads, decomposition + NCCL inside libshgpu) against the one-GPU run of the same script."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    for name, script in (("wall", open(os.path.join(ROOT, "examples", "in.wall_settle")).read()
                          .replace("run             5000", "run             800").replace("custom 5000", "custom 800")),
                         ("shear", open(os.path.join(ROOT, "examples", "in.shear_box")).read())):
        dumps = []
        for ng in (1, 2):
            d = tmp_path / ("%s_%d" % (name, ng)); d.mkdir()
            (d / "in.case").write_text(script)
            out = run_shlmp("in.case", d, extra=("-gpus", str(ng)))
            assert ("on %d GPU" % ng) in out
            dumps.append(read_last_dump(next(p for p in d.iterdir() if p.name.startswith("dump."))))
        a, b = dumps
        assert a.shape == b.shape and len(a) >= 320
        assert np.abs(a[:, 2:5] - b[:, 2:5]).max() < 1e-9, name
        assert np.abs(a[:, 5:9] - b[:, 5:9]).max() < 1e-9, name
        assert np.abs(a[:, 9:12] - b[:, 9:12]).max() < 1e-8, name
