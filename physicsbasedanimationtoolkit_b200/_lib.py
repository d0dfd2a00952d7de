"""ctypes loader for ``libvbdx.so`` -- the C-ABI of ``include/vbdx.h``.

There is no Python or CPU fallback: if the shared library is missing the import of anything
that needs it raises, and if no CUDA device is present ``vbdx_create`` fails with
``VBDX_NO_DEVICE``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VBDX_LIBRARY") or os.path.join(_HERE, "libvbdx.so")  # VBDX_LIBRARY: A/B runs against another build

VBDX_OK, VBDX_INVALID_ARGUMENT, VBDX_NO_DEVICE, VBDX_CUDA_ERROR, VBDX_OUT_OF_MEMORY, VBDX_UNSUPPORTED = range(6)
FLAG_ADAPTIVE_VBD_GPU_HISTORY = 1
FLAG_NATURAL_VERTEX_ORDER = 2
KERNEL_DEFAULT, KERNEL_DIRECT, KERNEL_TMA, KERNEL_PIPELINED = 0, 1, 2, 3


class DataDesc(C.Structure):
    """``vbdx_data_desc`` (include/vbdx.h)."""
    _fields_ = [
        ("abi_version", C.c_uint32), ("struct_size", C.c_uint32),
        ("nV", C.c_int64), ("nT", C.c_int64),
        ("X", C.c_void_p), ("E", C.c_void_p), ("v", C.c_void_p), ("aext", C.c_void_p),
        ("m", C.c_void_p), ("rhoe", C.c_void_p), ("lame", C.c_void_p),
        ("dbc", C.c_void_p), ("nDbc", C.c_int64), ("colors", C.c_void_p),
        ("ordering", C.c_int32), ("selection", C.c_int32), ("strategy", C.c_int32),
        ("acceleration", C.c_int32), ("omega_mode", C.c_int32), ("material", C.c_int32),
        ("kD", C.c_double), ("detHZero", C.c_double), ("rho", C.c_double),
        ("B", C.c_void_p), ("V", C.c_void_p), ("nCV", C.c_int64), ("F", C.c_void_p), ("nF", C.c_int64),
        ("muC", C.c_double), ("muF", C.c_double), ("epsv", C.c_double),
        ("active_set_update_frequency", C.c_int32),
        ("device", C.c_int32), ("tile_iters", C.c_int32), ("flags", C.c_int32),
        ("kernel_variant", C.c_int32), ("ring_slots", C.c_int32),
        ("ghosts", C.c_void_p), ("nGhosts", C.c_int64),
        ("consumer_warps", C.c_int32), ("window_size", C.c_int32),
        ("n_colors", C.c_int32), ("nesterov_start", C.c_int32), ("nesterov_L", C.c_double),
        ("tr_eta", C.c_double), ("tr_tau", C.c_double), ("tr_curved", C.c_int32), ("reserved0", C.c_int32),
    ]


class XpbdDesc(C.Structure):
    """``vbdx_xpbd_desc`` (include/vbdx.h)."""
    _fields_ = [
        ("abi_version", C.c_uint32), ("struct_size", C.c_uint32),
        ("nV", C.c_int64), ("nT", C.c_int64),
        ("X", C.c_void_p), ("T", C.c_void_p), ("v", C.c_void_p), ("aext", C.c_void_p), ("minv", C.c_void_p), ("lame", C.c_void_p),
        ("dbc", C.c_void_p), ("nDbc", C.c_int64),
        ("Pptr", C.c_void_p), ("Padj", C.c_void_p), ("nPartitions", C.c_int32), ("nClusterPartitions", C.c_int32),
        ("SGptr", C.c_void_p), ("SGadj", C.c_void_p), ("Cptr", C.c_void_p), ("Cadj", C.c_void_p),
        ("alphaSNH", C.c_void_p), ("betaSNH", C.c_void_p),
        ("BV", C.c_void_p), ("V", C.c_void_p), ("nCV", C.c_int64), ("F", C.c_void_p), ("nF", C.c_int64),
        ("muV", C.c_void_p), ("alphaC", C.c_void_p), ("betaC", C.c_void_p),
        ("muS", C.c_double), ("muD", C.c_double),
        ("active_set_update_frequency", C.c_int32), ("device", C.c_int32),
    ]


class Info(C.Structure):
    """``vbdx_info`` (include/vbdx.h)."""
    _fields_ = [
        ("nV", C.c_int64), ("nT", C.c_int64), ("nActiveVertices", C.c_int64),
        ("nIncidences", C.c_int64), ("nRecordSlots", C.c_int64),
        ("nColors", C.c_int32), ("nTiles", C.c_int32),
        ("gridBlocks", C.c_int32), ("blockThreads", C.c_int32),
        ("device", C.c_int32), ("smCount", C.c_int32),
        ("deviceBytes", C.c_int64), ("kernelLaunches", C.c_int64), ("lastStepMs", C.c_double),
        ("nRingEntries", C.c_int64), ("nGhosts", C.c_int64), ("nonFiniteVertices", C.c_int64),
    ]


# every symbol declared in include/vbdx.h: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "vbdx_data_desc_init": (None, [C.POINTER(DataDesc)]),
    "vbdx_create": (C.c_int, [C.POINTER(DataDesc), C.POINTER(_H)]),
    "vbdx_destroy": (C.c_int, [_H]),
    "vbdx_step": (C.c_int, [_H, C.c_double, C.c_int32, C.c_int32]),
    "vbdx_step_async": (C.c_int, [_H, C.c_double, C.c_int32, C.c_int32]),
    "vbdx_synchronize": (C.c_int, [_H]),
    "vbdx_step_partial": (C.c_int, [_H, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "vbdx_objective": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "vbdx_set_vertex_field": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "vbdx_get_vertex_field": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "vbdx_set_vertex_field_async": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "vbdx_get_vertex_field_async": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "vbdx_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "vbdx_host_free": (C.c_int, [C.c_void_p]),
    "vbdx_set_positions_f32": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_set_positions_f64": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_set_velocities_f32": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_set_velocities_f64": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_set_external_acceleration_f32": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_set_external_acceleration_f64": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_get_positions_f32": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_get_positions_f64": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_get_velocities_f32": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_get_velocities_f64": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_set_detH_zero": (C.c_int, [_H, C.c_double]),
    "vbdx_set_rayleigh_damping": (C.c_int, [_H, C.c_double]),
    "vbdx_set_initialization_strategy": (C.c_int, [_H, C.c_int32]),
    "vbdx_bvh_create": (C.c_int, [C.c_int64, C.POINTER(_H)]),
    "vbdx_bvh_destroy": (C.c_int, [_H]),
    "vbdx_bvh_build": (C.c_int, [_H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbdx_bvh_get": (C.c_int, [_H] + [C.c_void_p] * 8),
    "vbdx_bvh_detect_overlaps": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vbdx_bvh_nearest_triangles": (C.c_int, [_H, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vbdx_contact_create": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(_H)]),
    "vbdx_contact_destroy": (C.c_int, [_H]),
    "vbdx_contact_initialize_active_set": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbdx_contact_update_active_set": (C.c_int, [_H, C.c_void_p]),
    "vbdx_contact_finalize_active_set": (C.c_int, [_H, C.c_void_p]),
    "vbdx_contact_set_eps": (C.c_int, [_H, C.c_float]),
    "vbdx_contact_get": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbdx_debug_plan_create": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(_H)]),
    "vbdx_debug_plan_get": (C.c_int, [_H, C.c_int32, C.c_void_p]),
    "vbdx_debug_plan_destroy": (C.c_int, [_H]),
    "vbdx_create_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "vbdx_batch_offsets": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "vbdx_set_block_size": (C.c_int, [_H, C.c_int32]),
    "vbdx_set_line_search_guard": (C.c_int, [_H, C.c_int32]),
    "vbdx_set_scene_bounding_box": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "vbdx_set_stream": (C.c_int, [_H, C.c_void_p]),
    "vbdx_get_internal_ids": (C.c_int, [_H, C.c_void_p]),
    "vbdx_dist_ipc_handles": (C.c_int, [_H, C.c_void_p]),
    "vbdx_dist_connect": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "vbdx_dist_stats": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "vbdx_get_contact_state": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbdx_debug_bvh_build": (C.c_int, [C.c_int64] + [C.c_void_p] * 11),
    "vbdx_debug_trace": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_int64]),
    "vbdx_get_info": (C.c_int, [_H, C.POINTER(Info)]),
    "vbdx_get_adjacency": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbdx_get_element_data": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbdx_get_colors": (C.c_int, [_H, C.c_void_p]),
    "vbdx_greedy_color": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vbdx_greedy_color_device": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vbdx_debug_contact_pairs": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p]),
    "vbdx_debug_contact_penalties": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "vbdx_xpbd_desc_init": (None, [C.POINTER(XpbdDesc)]),
    "vbdx_xpbd_create": (C.c_int, [C.POINTER(XpbdDesc), C.POINTER(_H)]),
    "vbdx_xpbd_destroy": (C.c_int, [_H]),
    "vbdx_xpbd_step": (C.c_int, [_H, C.c_double, C.c_int32, C.c_int32]),
    "vbdx_xpbd_set_positions": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_xpbd_set_velocities": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_xpbd_set_external_acceleration": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_xpbd_get_positions": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_xpbd_get_velocities": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "vbdx_xpbd_set_compliance": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_int64]),
    "vbdx_xpbd_set_friction_coefficients": (C.c_int, [_H, C.c_double, C.c_double]),
    "vbdx_xpbd_set_scene_bounding_box": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "vbdx_xpbd_get_info": (C.c_int, [_H, C.c_void_p]),
    "vbdx_xpbd_get_contact_state": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbdx_graph_greedy_color": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vbdx_last_error": (C.c_char_p, []),
    "vbdx_abi_version": (C.c_int32, []),
    "vbdx_device_count": (C.c_int32, []),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded library; raises ``ImportError`` when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C physicsbasedanimationtoolkit_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    """Translate a ``vbdx_status`` into the exception the reference would raise."""
    if status == VBDX_OK:
        return
    msg = lib().vbdx_last_error().decode()
    if status == VBDX_INVALID_ARGUMENT:
        raise ValueError(msg)  # nanobind maps std::invalid_argument to ValueError
    if status == VBDX_OUT_OF_MEMORY:
        raise MemoryError(msg)
    if status == VBDX_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
