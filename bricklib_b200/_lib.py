"""ctypes binding of libbrick_b200.so (include/bricklib_b200.h).  Loading never touches the GPU; every compute
entry point fails with BK_ECUDA when no device is present -- there is no CPU fallback in this package."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbrick_b200.so")

BK_OK = 0
BK_EUNSUPPORTED = -4
ST_7PT, ST_MPI7PT, ST_MPI13PT, ST_MPI25PT, ST_MPI125PT = range(5)
STENCILS = {"7pt": 0, "mpi7pt": 1, "mpi13pt": 2, "mpi25pt": 3, "mpi125pt": 4}
KERNEL_AUTO, KERNEL_BRICK, KERNEL_TILED = 0, 1, 2
PART_ALL, PART_READY, PART_REST, PART_THIN, PART_GRID_TOPOLOGY = 0, 1, 2, 4, 8
FUSED_STAGED, FUSED_COMPOSED, FUSED_COMPOSED_WIDE = 0, 1, 2   # bk_stencil_fused_variant_set: k_star2 / the composed diamond (4x4 | 8x4 tiles)
IPC_HANDLE_BYTES = 64

vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
ip = C.POINTER(C.c_int)
up = C.POINTER(C.c_uint)
lp = C.POINTER(C.c_long)
dp = C.POINTER(C.c_double)


class Region(C.Structure):
    _fields_ = [("neighbor", u64), ("skin_st", C.c_uint), ("skin_ed", C.c_uint), ("pos", C.c_uint), ("len", C.c_uint),
                ("first_pad", C.c_uint), ("last_pad", C.c_uint)]


class Field(C.Structure):
    _fields_ = [("adj", vp), ("inp", vp), ("in_step", sz), ("out", vp), ("out_step", sz)]


class Seg(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("bytes", sz)]


class Box(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("n", C.c_long * 3), ("src_stride", C.c_long * 2), ("dst_stride", C.c_long * 2)]


class Pointwise(C.Structure):
    _fields_ = [("op", C.c_int), ("c", C.c_double)]


POINTWISE_OPS = {"max": 1, "min": 2, "abs": 3}


class StitchBox(C.Structure):
    _fields_ = [("first", C.c_ulong), ("count", C.c_ulong), ("lo", C.c_ulong * 3), ("n", C.c_long * 3),
                ("wrap", C.c_int * 3), ("is_box", C.c_int)]


class Tap(C.Structure):
    _fields_ = [("di", C.c_int), ("dj", C.c_int), ("dk", C.c_int), ("c", C.c_double)]


# name -> (restype, argtypes); the list doubles as the export check in tests/test_abi.py
SIGNATURES = {
    "bk_version": (C.c_char_p, []),
    "bk_last_error": (C.c_char_p, []),
    "bk_stencil_radius": (C.c_int, [C.c_int]),
    "bk_stencil_st_iter": (C.c_int, [C.c_int]),
    "bk_stencil_points": (C.c_int, [C.c_int]),
    "bk_stencil_fused_steps": (C.c_int, [C.c_int]),
    "bk_stencil_fused_variant_set": (C.c_int, [C.c_int]),
    "bk_stencil_fused_variant_get": (C.c_int, []),
    "bk_device_count": (C.c_int, [ip]),
    "bk_set_device": (C.c_int, [C.c_int]),
    "bk_bind_host_to_device": (C.c_int, []),
    "bk_dev_alloc": (C.c_int, [C.POINTER(vp), sz]),
    "bk_dev_free": (C.c_int, [vp]),
    "bk_dev_memset": (C.c_int, [vp, C.c_int, sz, vp]),
    "bk_host_alloc": (C.c_int, [C.POINTER(vp), sz]),
    "bk_host_free": (C.c_int, [vp]),
    "bk_memcpy_h2d": (C.c_int, [vp, vp, sz, vp]),
    "bk_memcpy_d2h": (C.c_int, [vp, vp, sz, vp]),
    "bk_memcpy_d2d": (C.c_int, [vp, vp, sz, vp]),
    "bk_stream_create": (C.c_int, [C.POINTER(vp)]),
    "bk_stream_create_priority": (C.c_int, [C.POINTER(vp), C.c_int]),
    "bk_stream_destroy": (C.c_int, [vp]),
    "bk_stream_sync": (C.c_int, [vp]),
    "bk_device_sync": (C.c_int, []),
    "bk_event_create": (C.c_int, [C.POINTER(vp)]),
    "bk_event_destroy": (C.c_int, [vp]),
    "bk_event_record": (C.c_int, [vp, vp]),
    "bk_event_sync": (C.c_int, [vp]),
    "bk_event_elapsed_ms": (C.c_int, [vp, vp, C.POINTER(C.c_float)]),
    "bk_stream_wait_event": (C.c_int, [vp, vp]),
    "bk_init_grid": (C.c_int, [lp, up, up]),
    "bk_decomp_create": (C.c_int, [C.POINTER(vp), up, C.c_uint]),
    "bk_decomp_destroy": (C.c_int, [vp]),
    "bk_decomp_nbricks": (C.c_uint, [vp]),
    "bk_decomp_sep_pos": (C.c_int, [vp, up]),
    "bk_decomp_tdims": (C.c_int, [vp, up]),
    "bk_decomp_grid": (up, [vp]),
    "bk_decomp_adj": (up, [vp]),
    "bk_decomp_nregions": (C.c_int, [vp]),
    "bk_decomp_region": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(Region)]),
    "bk_decomp_skin_size": (C.c_int, [vp, lp]),
    "bk_decomp_list": (C.c_long, [vp, C.c_int, up]),
    "bk_rank_map": (C.c_int, [ip, ip, C.POINTER(u64), ip]),
    "bk_zmort_encode": (C.c_ulong, [C.POINTER(C.c_ulong)]),
    "bk_zmort_decode": (C.c_int, [C.c_ulong, C.POINTER(C.c_ulong)]),
    "bk_stitch_box": (C.c_int, [C.c_ulong, C.c_ulong, C.c_ulong, C.POINTER(StitchBox)]),
    "bk_stitch_dims": (C.c_int, [vp, C.POINTER(StitchBox), up]),
    "bk_stitch_grid": (C.c_int, [vp, C.POINTER(StitchBox), up]),
    "bk_stitch_region_needed": (C.c_int, [vp, C.POINTER(StitchBox), C.c_ulong, C.c_int]),
    "bk_copy_to_brick": (C.c_int, [lp, lp, lp, vp, vp, vp, sz, vp]),
    "bk_copy_from_brick": (C.c_int, [lp, lp, lp, vp, vp, vp, sz, vp]),
    "bk_compare_brick": (C.c_int, [lp, lp, lp, vp, vp, vp, sz, C.c_double, C.POINTER(C.c_ulonglong), dp, vp]),
    "bk_fill_synthetic": (C.c_int, [vp, up, lp, lp, u64, vp, sz, vp]),
    "bk_synthetic_value": (C.c_double, [u64, u64]),
    "bk_compare_storage": (C.c_int, [vp, up, up, up, vp, sz, vp, sz, C.c_double, C.POINTER(C.c_ulonglong), dp, vp]),
    "bk_stencil_apply": (C.c_int, [C.c_int, C.POINTER(Field), vp, up, up, up, dp, C.c_uint, vp]),
    "bk_adjacency_forget": (C.c_int, [vp]),
    "bk_stencil_apply_part": (C.c_int, [C.c_int, C.POINTER(Field), vp, up, up, up, dp, up, up, C.c_int, vp]),
    "bk_stencil_advance": (C.c_int, [C.c_int, C.c_int, C.POINTER(Field), vp, up, up, up, dp, up, up, C.c_int, vp]),
    "bk_stencil_advance_remote": (C.c_int, [C.c_int, C.c_int, C.POINTER(Field), vp, up, up, up, dp, up, up, C.c_int, vp, C.c_uint,
                                            C.c_uint, vp]),
    "bk_stencil_apply_list": (C.c_int, [C.c_int, C.POINTER(Field), vp, sz, dp, vp]),
    "bk_array_stencil_apply": (C.c_int, [C.c_int, vp, vp, lp, lp, lp, dp, vp]),
    "bk_stencil_apply_multi": (C.c_int, [C.c_int, vp, C.c_uint, vp, up, up, up, dp, vp]),
    "bk_stencil_compile": (C.c_int, [C.POINTER(vp), C.POINTER(Tap), C.c_int]),
    "bk_stencil_compile_pointwise": (C.c_int, [C.POINTER(vp), C.POINTER(Tap), C.c_int, C.POINTER(Pointwise),
                                               C.POINTER(Pointwise)]),
    "bk_stencil_def_destroy": (C.c_int, [vp]),
    "bk_stencil_def_source": (C.c_int, [vp, C.c_char_p, sz, C.POINTER(sz)]),
    "bk_stencil_def_info": (C.c_int, [vp, ip, ip, ip, ip, ip]),
    "bk_stencil_def_apply": (C.c_int, [vp, C.POINTER(Field), vp, up, up, up, C.c_uint, vp]),
    "bk_stencil_def_advance": (C.c_int, [vp, C.c_int, C.POINTER(Field), vp, up, up, up, up, up, C.c_int, C.c_uint, vp]),
    "bk_launch_count": (C.c_ulonglong, []),
    "bk_xplan_create": (C.c_int, [C.POINTER(vp), C.POINTER(Seg), C.c_int]),
    "bk_xplan_create_boxes": (C.c_int, [C.POINTER(vp), C.POINTER(Box), C.c_int]),
    "bk_xplan_destroy": (C.c_int, [vp]),
    "bk_xplan_bytes": (sz, [vp]),
    "bk_xplan_set_shape": (C.c_int, [vp, C.c_int, C.c_int]),
    "bk_xplan_run": (C.c_int, [vp, vp]),
    "bk_xplan_run_sync": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, u64, vp]),
    "bk_xplan_run_gate": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, vp, u64, vp]),
    "bk_xplan_run_ce": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, u64, vp]),
    "bk_flags_signal": (C.c_int, [C.POINTER(vp), C.c_int, u64, vp]),
    "bk_flags_wait": (C.c_int, [C.POINTER(vp), C.c_int, u64, vp]),
    "bk_ipc_export": (C.c_int, [vp, C.c_char_p]),
    "bk_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "bk_ipc_close": (C.c_int, [vp]),
    "bk_peer_enable": (C.c_int, [C.c_int]),
}


class BrickError(RuntimeError):
    pass


_LIB = None


def load():
    """dlopen the in-tree library (built by __graft_entry__.build() / bricklib_b200/csrc/Makefile)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise BrickError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(rc):
    if rc != BK_OK:
        raise BrickError(f"libbrick_b200 error {rc}: {load().bk_last_error().decode()}")
    return rc
