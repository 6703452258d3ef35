"""ctypes binding of libcellseg_sm100.so (C ABI in include/cellseg_b200.h).

The CUDA library is the only compute path of the package.  `load()` raises if
it has not been built — there is no CPU fallback and no lazy JIT.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t

# name -> (restype, argtypes); kept in step with include/cellseg_b200.h
SIGNATURES = {
    "cs_last_error": (C.c_char_p, []),
    "cs_abi_version": (_i, []),
    "cs_kernel_launch_count": (C.c_ulonglong, []),
    "cs_profile_enable": (_i, [_i]),
    "cs_profile_read": (_i, [_i, C.c_char_p, _sz, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "cs_postprocess_workspace_bytes": (_sz, [_i, _i, _i]),
    "cs_postprocess": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "cs_postprocess_host": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "cs_postprocess_packed_slot": (_vp, [_vp, _sz, _i, _i, _i]),
    "cs_postprocess_packed": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "cs_scale_lut_u8": (_i, [_vp, _i, _sz, _vp, _vp, _vp]),
    "cs_u8hwc_to_f32chw": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "cs_prenorm_hist": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "cs_prenorm_apply": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "cs_nhwc_stat_chunks": (_i, [_i, _i]),
    "cs_nhwc_stats": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "cs_nhwc_norm_prelu": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, C.c_float, _vp]),
    "cs_instnorm_prelu_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, C.c_float, _vp, _vp, _i, _i, _vp]),
    "cs_instnorm_prelu_nhwc_slices": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, C.c_float, _vp, _vp, _i, _i, _vp]),
    "cs_instnorm_prelu_nhwc_shuffle": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_float, _vp, _vp]),
    "cs_unet_stem": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cs_unet_stem_u8": (_i, [_vp, _vp] + [_i] * 8 + [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cs_unet_head_partial_chunks": (_i, [_i, _i]),
    "cs_unet_head_convt": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "cs_unet_head_final": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, C.c_float, _vp, _vp]),
    "cs_instnorm_prelu": (_i, [_vp, _vp, _vp, _vp, _vp, _i, C.c_longlong, _i, C.c_float, _vp]),
    "cs_interior_boundary": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "cs_interior_combine": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "cs_u8_max": (_i, [_vp, _i, _sz, _vp, _vp]),
    "cs_scale_lut_from_max": (_i, [_vp, _i, _vp, _vp]),
    "cs_strip_workspace_bytes": (_sz, [_i, _i]),
    "cs_strip_packed_slot": (_vp, [_vp, _sz, _i, _i]),
    "cs_strip_begin": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "cs_strip_filter_export": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "cs_strip_filter_resume": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cs_strip_label_export": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cs_strip_resolve_workspace_bytes": (_sz, [C.c_longlong]),
    "cs_strip_label_resolve": (_i, [_vp, _vp, _i, _i, _i, _i, C.c_int32, C.c_int32, C.c_longlong, _vp, _vp, _vp, _sz, _vp]),
    "cs_strip_final": (_i, [_i, _i, _vp, _vp, _vp, _sz, _vp]),
    "cs_gather_windows": (_i, [_vp] + [_i] * 8 + [_vp, _vp, _i, _i, _vp, _vp]),
    "cs_gather_windows_u8": (_i, [_vp, _vp] + [_i] * 8 + [_vp, _vp, _i, _i, _vp, _vp]),
    "cs_blend_softmax_threshold": (_i, [_vp] + [_i] * 9 + [_vp, _vp, _vp, _vp, _vp, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "cs_label": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "cs_label_host": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "cs_remove_small": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "cs_certify_large": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "cs_remove_small_host": (_i, [_vp, _vp, _i, _i, _i, _i, _i]),
    "cs_label_max": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "cs_relabel_workspace_bytes": (_sz, [C.c_int32]),
    "cs_relabel": (_i, [_vp, _i, _i, _i, _i, C.c_int32, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "cs_remove_boundary_cells_host": (_i, [_vp, _i, _i, _vp, _vp]),
    "cs_relabel_sequential_host": (_i, [_vp, _i, _i, _vp, _vp]),
    "cs_overlap_workspace_bytes": (_sz, [C.c_uint32]),
    "cs_label_overlap_sparse": (_i, [_vp, _i, _vp, _i, _i, _i, C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp, _vp,
                                     C.c_uint32, _vp, _sz, _vp]),
    "cs_overlap_dense": (_i, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_int32, C.c_int32, _vp, _vp]),
    "cs_iou_dense": (_i, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]),
    "cs_iou_edges": (_i, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_int32, C.c_int32, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cs_count_nonzero3": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "cs_count_nonzero3_host": (_i, [_vp, _vp, _sz, _vp]),
    "cs_label_overlap_host": (_i, [_vp, _vp, _i, _i, C.c_int32, C.c_int32, _vp]),
}


class CellSegError(RuntimeError):
    pass


def bind(lib: C.CDLL) -> C.CDLL:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CellSegError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). cellseg_b200 has no CPU fallback.")
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def check(lib: C.CDLL, rc: int) -> None:
    if rc != 0:
        raise CellSegError(f"libcellseg status {rc}: {lib.cs_last_error().decode(errors='replace')}")


def profile_enable(on: bool = True) -> None:
    """Switch the library's per-kernel CUDA-event timing on (clearing the accumulators) or off."""
    load().cs_profile_enable(int(on))


def profile_read() -> dict:
    """{kernel name: (total_ms, launches)} recorded since profile_enable(True); waits for the recorded events."""
    lib = load()
    out, i = {}, 0
    name = C.create_string_buffer(128)
    ms, n = C.c_double(), C.c_ulonglong()
    while lib.cs_profile_read(i, name, 128, C.byref(ms), C.byref(n)):
        if n.value:
            out[name.value.decode()] = (ms.value, n.value)
        i += 1
    return out
