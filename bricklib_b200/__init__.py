"""bricklib_b200 -- B200-native hot path of bricklib: FP64 brick stencils + ghost-zone exchange.

The product is libbrick_b200.so (CUDA kernels for sm_100a behind the C ABI in include/bricklib_b200.h).  This package is
the thin host-side mirror of the reference interface used by tests/, bench.py and the weak-scaling loop; the C++
template surface lives in include/*.h and the drivers in drivers/.
"""
from ._lib import (BK_OK, FUSED_COMPOSED, FUSED_COMPOSED_WIDE, FUSED_STAGED, KERNEL_AUTO, KERNEL_BRICK, KERNEL_TILED, PART_ALL, PART_READY, PART_REST, PART_THIN, STENCILS,  # noqa: F401
                   BrickError, load)
from .core import (BRICK, Brick, BrickDecomp, BrickInfo, BrickStorage, DeviceBuffer, DeviceGrid, Event,  # noqa: F401
                   ExchangeView, ArrayExchangeView, array_stencil, bitset_of, StitchedGrid, compareBrick, copyFromBrick, copyToBrick, device_sync, init_grid, stencil,
                   stencil_advance, stencil_list, stencil_part, fill_synthetic, synthetic_field, compare_storage, section_owner, section_range, strong_pull_plan, zmort_decode, zmort_encode, Unsupported)
from .weak import ArrayDomain, FieldPipeline, WeakDomain, shell_boxes  # noqa: F401


def have_gpu():
    import ctypes
    n = ctypes.c_int()
    return load().bk_device_count(ctypes.byref(n)) == 0 and n.value > 0


def fused_variant(variant=None):
    """the kernel behind stencil_advance(steps=2) for radius-1 stars: FUSED_STAGED (k_star2), FUSED_COMPOSED or
    FUSED_COMPOSED_WIDE (the composed 25-point diamond on 4x4- / 8x4-brick tiles); sets it when given, returns the previous / current value"""
    L = load()
    if variant is None:
        return L.bk_stencil_fused_variant_get()
    from ._lib import check
    before = L.bk_stencil_fused_variant_set(int(variant))
    check(min(before, 0))
    return before


from .dsl import CompiledStencil, LoweringError, compile_stencil, lower  # noqa: E402,F401
