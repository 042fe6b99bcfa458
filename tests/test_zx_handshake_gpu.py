"""The device-side flag handshake that orders the ranks of the one-process-per-GPU weak loop, on ONE GPU (the driver's
test box has one): tests/handshake_case.py runs 2 or 4 emulated ranks concurrently on their own streams, ordered by the
flags alone, and compares with the lock-step loop and the oracle.  In a child process under a timeout: a protocol bug
shows as a hang, and a hang must not take the suite with it.  (The CUDA-IPC mapping itself needs two processes:
tests/test_multi_gpu.py, N >= 2 boxes.)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ranks,extra", [(2, []), (4, []), (2, ["--no-overlap"])])
def test_flag_handshake_orders_concurrent_ranks_on_one_gpu(ranks, extra):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "handshake_case.py"), "--ranks", str(ranks), *extra]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    except subprocess.TimeoutExpired:
        pytest.fail("the handshake case did not finish in 240 s (a rank waits for a flag nobody raises)")
    assert r.returncode == 0 and "handshake ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
