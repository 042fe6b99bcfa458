#!/usr/bin/env python3
"""What the host <-> device links of this box can carry: plain cudaMemcpyAsync of 1 GiB pinned buffers, H2D alone, D2H
alone and both directions at once, on every rank of a torchrun launch at the same time (or on one GPU without torchrun).
The end-to-end leg of bench.py moves 1.07 GB up and 1.07 GB down per GPU per step: this is its ceiling.
   python tools/pcie_probe.py            |  python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import bricklib_b200 as bk

rank, world, dist, _ = bench.dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
L = bk.load()
ck = bk._lib.check
if world == 1:
    ck(L.bk_set_device(0))
bound = L.bk_bind_host_to_device()
n = 1 << 30
h_in, h_out = C.c_void_p(), C.c_void_p()
ck(L.bk_host_alloc(C.byref(h_in), n)); ck(L.bk_host_alloc(C.byref(h_out), n))
C.memset(h_in, 1, n); C.memset(h_out, 0, n)
d_a, d_b = bk.DeviceBuffer(n), bk.DeviceBuffer(n)
s_up, s_down = C.c_void_p(), C.c_void_p()
ck(L.bk_stream_create(C.byref(s_up))); ck(L.bk_stream_create(C.byref(s_down)))

def timed(up, down, reps=4):
    bench.barrier(dist)
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            ck(L.bk_memcpy_h2d(d_a.ptr, h_in, n, s_up))
        if down:
            ck(L.bk_memcpy_d2h(h_out, d_b.ptr, n, s_down))
    ck(L.bk_stream_sync(s_up)); ck(L.bk_stream_sync(s_down))
    return bench.max_over_ranks(dist, (time.perf_counter() - t0) / reps)

timed(True, True, 1)
res = {"gpus": world, "numa_bound": bound == 0, "GiB": 1,
       "h2d_GB/s_per_gpu": n / timed(True, False) / 1e9, "d2h_GB/s_per_gpu": n / timed(False, True) / 1e9}
both = timed(True, True)
res["duplex_GB/s_per_gpu_per_direction"] = n / both / 1e9
res["duplex_GB/s_aggregate_per_direction"] = world * n / both / 1e9
if rank == 0:
    print(json.dumps(res))
if dist is not None:
    bench.barrier(dist)
    dist.destroy_process_group()
