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
                                           ("in.wall_settle", "run 5000 with 1000 atoms, 1 shape(s), 1 wall(s) OK"),
                                           ("in.shear_box", "run 300 with 320 atoms, 1 shape(s), 0 wall(s) OK")])
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
    (HEAD + "pair_style spherharm\npair_coeff 1 1 1000 1 0.5\n", "Incorrect args for pair coefficients"),
    (HEAD + "pair_style spherharm\npair_coeff 1 1 1000 1 -1 0 0\n", "Incorrect args for pair coefficients"),
    (HEAD + "fix 2 all deform 1 xz erate 0.1 remap v\n", "Only fix deform N xy erate <rate> remap v is supported"),
    (HEAD + "boundary p f p\npair_style spherharm\nfix 2 all deform 1 xy erate 0.1 remap v\nrun 1\n", "fix deform xy needs a box periodic in x and y"),
])
def test_malformed_scripts_are_rejected(text, msg):
    res = shlmp(script_text=text)
    assert res.returncode == 1
    assert "ERROR: " + msg in res.stderr, res.stderr


def test_dissipative_pair_coeff_and_gpus_flag_are_accepted():
    res = shlmp(script_text=HEAD + "pair_style spherharm\npair_coeff 1 1 1000 1 3.0 2.0 0.4\ncreate_atoms 1 single 0 0 0\nrun 5\n")
    assert res.returncode == 0 and "run 5 with 1 atoms" in res.stdout, res.stderr
    from lammps_spherharm_b200 import build as b
    r2 = subprocess.run([b.build_host(), "-check", "-gpus", "4", "-in", "in.shear_box"], cwd=EX, capture_output=True, text=True, timeout=120)
    assert r2.returncode == 0 and "OK" in r2.stdout, r2.stderr
