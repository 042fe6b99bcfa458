"""The exchange inside the sweep (bk_stencil_advance_remote: the first pass of a period reads the ghost bricks in place from
the neighbours' storages instead of from a pulled copy) against the pull, on one GPU, where the neighbours of the periodic
single-rank domain are the domain itself.  Through tools/direct_exchange_trial.py in a child process: the in-place kernels
(bk_stencil_remote.cu) were written in a round without GPU time -- their ordering and bookkeeping are covered on the CPU
stand-in (tests/test_hostdev.py), their first run on hardware must not be able to take the suite with it."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def trial():
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "direct_exchange_trial.py"), "--size", "128", "--periods", "3"],
                           capture_output=True, text=True, timeout=300, cwd=ROOT)
    except subprocess.TimeoutExpired:
        pytest.fail("the trial did not finish in 300 s")
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("name", ["mpi13pt", "mpi25pt", "mpi125pt"])
def test_reading_ghost_bricks_in_place_gives_the_same_field_as_pulling_them(trial, name):
    assert trial[name].get("ok") and trial[name]["mismatches"] == 0, trial[name]


def test_the_composed_two_step_kernel_reads_ghost_bricks_in_place_too(trial):
    """7-point: two steps per pass -- only the composed kernel has an in-place variant (the staged one keeps the pull)"""
    assert trial["mpi7pt"].get("ok") and trial["mpi7pt"]["mismatches"] == 0, trial["mpi7pt"]
