"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from this repo's CPU oracle; PARITY UNPINNED:
the reference mount has no vectors).  CPU: the oracle still reproduces them bit for bit (its arithmetic has not drifted).
GPU: the CUDA path, through the C ABI, against the committed numbers, without the oracle at run time."""
import importlib.util
import os
import sys

import numpy as np
import pytest

import shpkg

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)
pkg = shpkg.load()
W = pkg.workloads
CASES = MG.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    cfg, nsteps = CASES[name]
    got = MG.run_case(cfg, nsteps)
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    for k in ref.files:
        assert np.array_equal(np.asarray(got[k]), ref[k]), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 16])
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_golden(name, variant):
    cfg, nsteps = CASES[name]
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    g = pkg.ShGpu(); W.apply(g, cfg); g.set_pair_tuning(0, 0, variant)
    if nsteps:
        g.run(nsteps)
    else:
        g.compute_forces()
    p = g.get_pairs()
    nl = g.dd_info()["nlocal"]
    tag = g.get_tags()[:nl]
    a = g.get_atoms()
    o = np.argsort(tag)
    f, tq = a["f"][:nl][o], a["torque"][:nl][o]
    fs = max(np.abs(ref["f"]).max(), 1e-300)
    tol = 1e-10 if not nsteps else 1e-7          # a 40-step trajectory amplifies rounding differences
    assert np.abs(f - ref["f"]).max() <= tol * fs and np.abs(tq - ref["torque"]).max() <= tol * fs
    if not cfg.get("shear"):                     # (a sheared box lists the pairs of ghost images under the image's tag pair twice)
        key = {(int(i), int(j)): k for k, (i, j) in enumerate(zip(p["tag_i"], p["tag_j"]))}
        idx = np.array([key[(int(i), int(j))] for i, j in zip(ref["tag_i"], ref["tag_j"])])
        assert len(key) == len(ref["V"])
        V = p["V"][idx]
        assert np.array_equal(V > 0, ref["V"] > 0)
        m = ref["V"] > 0
        assert np.abs(V[m] - ref["V"][m]).max() <= 1e-10 * ref["V"][m].max() if m.any() else True
        fn = np.linalg.norm(ref["F"][m], axis=1)
        assert (np.linalg.norm(p["F"][idx][m] - ref["F"][m], axis=1) <= 1e-10 * np.maximum(fn, 1e-300)).all()
        assert int(g.get_counters()["nodes_inside"]) == int(ref["nodes_inside"])
    g.close()
