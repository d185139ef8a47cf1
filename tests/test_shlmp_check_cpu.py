"""CPU tests of the shlmp input-script surface in -check mode (parse + validate, no device): the example scripts
are accepted, and malformed scripts fail with LAMMPS-style messages."""
import os
import subprocess

import pytest

import shpkg

pkg = shpkg.load()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")


def shlmp(script_text=None, script_file=None, cwd=EX):
    from lammps_spherharm_b200 import build as b
    exe = b.build_host()
    if script_file is None:
        res = subprocess.run([exe, "-check"], input=script_text, cwd=cwd, capture_output=True, text=True, timeout=120)
    else:
        res = subprocess.run([exe, "-check", "-in", script_file], cwd=cwd, capture_output=True, text=True, timeout=120)
    return res


@pytest.mark.parametrize("script,expect", [("in.two_particle", "run 1200 with 2 atoms, 1 shape(s), 0 wall(s) OK"),
                                           ("in.wall_settle", "run 5000 with 1000 atoms, 1 shape(s), 1 wall(s) OK")])
def test_examples_parse(script, expect):
    res = shlmp(script_file=script)
    assert res.returncode == 0, res.stderr
    assert expect in res.stdout


HEAD = "atom_style spherharm 20 32 64 ellipsoid_l20.sh\nregion box block -5 5 -5 5 -5 5\ncreate_box 1 box\n"


@pytest.mark.parametrize("text,msg", [
    ("pair_style lj/cut 2.5\n", "Unknown pair style lj/cut"),
    ("atom_style sphere\n", "Unknown atom style sphere"),
    ("atom_style spherharm 20 32\n", "Illegal atom_style spherharm command"),
    (HEAD + "pair_coeff 1 1 1000 1\n", "Pair_coeff command before pair_style is defined"),
    (HEAD + "pair_style spherharm\npair_coeff 2 2 1000 1\n", "Incorrect args for pair coefficients"),
    (HEAD + "fix 1 all langevin 1 1 1 1\n", "Unknown fix style langevin"),
    (HEAD + "create_atoms 3 single 0 0 0\npair_style spherharm\nrun 1\n", "Invalid atom type"),
    ("atom_style spherharm 20 32 64 nosuchfile.sh\nregion b block 0 1 0 1 0 1\ncreate_box 1 b\npair_style spherharm\nrun 1\n",
     "Cannot open shape file nosuchfile.sh"),
    (HEAD + "pair_style spherharm\nfoo bar\n", "Unknown command: foo"),
    ("atom_style spherharm 20 32 64 ellipsoid_l20.sh\npair_style spherharm\nrun 1\n", "Box must be defined before run"),
    (HEAD + "timestep -1\n", "Illegal timestep command"),
])
def test_malformed_scripts_are_rejected(text, msg):
    res = shlmp(script_text=text)
    assert res.returncode == 1
    assert "ERROR: " + msg in res.stderr, res.stderr
