"""One process per GPU over CUDA IPC / NVLink (the path `bench.py --gpus N` times): parity with the oracle's periodic
global sweep at every process-grid shape the box can hold.  Skipped on a single-GPU box (the N>1 host logic is covered on
CPU by tests/test_multirank_gloo.py and, on one GPU, by the lock-step multi-rank tests in test_gpu_stencil.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest

import bricklib_b200 as bk

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import ctypes as C
    n = C.c_int(0)
    try:
        bk.load().bk_device_count(C.byref(n))
    except Exception:
        return 0
    return n.value


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("transport", ["kernel", "ce"])  # one pull kernel / faces on the copy engines + narrow kernel
@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_weak_loop_one_process_per_gpu(n, transport):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_weak_check.py"), "--size",
           "32", "--periods", "2", "--transport", transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, BK_CE_MIN_BYTES="4096"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [x for x in r.stdout.splitlines() if x.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["max_rel"] < 1e-12, res
