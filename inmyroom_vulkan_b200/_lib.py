"""ctypes loader for libimrcd.so (the C ABI declared in include/imrcd.h).

There is no CPU fallback: if the shared library has not been built, importing a symbol raises
ImportError with the build command; if no sm_100 GPU is usable, imrcd_create fails with
IMRCD_E_NODEVICE and the Python wrapper raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libimrcd.so")

IMRCD_BUILD_MORTON = 0
IMRCD_BUILD_REFERENCE = 1

ERRORS = {0: "ok", -1: "cuda", -2: "argument", -3: "no sm_100 device (the library has no CPU path)", -4: "capacity", -5: "state"}


class EntityPair(C.Structure):
    _fields_ = [("entry_first", C.c_uint32), ("entry_second", C.c_uint32), ("entity_first", C.c_uint32),
                ("entity_second", C.c_uint32), ("n_hits", C.c_uint32), ("n_rays_first", C.c_uint32),
                ("n_rays_second", C.c_uint32), ("flags", C.c_uint32), ("avg_first", C.c_float * 3),
                ("avg_second", C.c_float * 3), ("delta_first", C.c_float * 3), ("delta_second", C.c_float * 3)]


class TriHit(C.Structure):
    _fields_ = [("pair", C.c_uint32), ("tri_first", C.c_uint32), ("tri_second", C.c_uint32), ("source", C.c_float * 3),
                ("target", C.c_float * 3), ("weight", C.c_float)]


class FrameStats(C.Structure):
    _fields_ = [("n_entries", C.c_uint64), ("n_pairs", C.c_uint64), ("n_sat_tests", C.c_uint64), ("n_combos", C.c_uint64),
                ("n_tri_tests", C.c_uint64), ("n_hits", C.c_uint64), ("n_coplanar_hits", C.c_uint64), ("n_colliding", C.c_uint64),
                ("traverse_launches", C.c_uint64), ("total_launches", C.c_uint64), ("n_queue_items", C.c_uint64), ("n_warp_iterations", C.c_uint64), ("trav_busy_cycles", C.c_uint64), ("trav_idle_polls", C.c_uint64), ("ms_total", C.c_float), ("ms_broad", C.c_float),
                ("ms_pair_setup", C.c_float), ("ms_traverse", C.c_float), ("ms_narrow", C.c_float), ("ms_reduce", C.c_float),
                ("n_contact_pairs", C.c_uint64), ("n_rays", C.c_uint64),
                ("n_rays_shot", C.c_uint64), ("n_responses", C.c_uint64), ("ms_response", C.c_float),
                ("n_merged", C.c_uint64), ("n_entries_local", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/imrcd.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SIGS = [
    ("imrcd_create", C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    ("imrcd_destroy", None, [_P]),
    ("imrcd_last_error", C.c_char_p, [_P]),
    ("imrcd_version", C.c_char_p, []),
    ("imrcd_abi_layout", C.c_int, [C.POINTER(C.c_uint64), C.c_uint64]),
    ("imrcd_mesh_create", C.c_int, [_P, _P, _P, _P, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("imrcd_gltf_open", C.c_int, [C.c_char_p, C.POINTER(_P), C.c_char_p, C.c_uint64]),
    ("imrcd_gltf_close", None, [_P]),
    ("imrcd_gltf_mesh_count", C.c_int, [_P, C.POINTER(C.c_uint32)]),
    ("imrcd_gltf_primitive_count", C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("imrcd_gltf_primitive", C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    ("imrcd_gltf_build_mesh", C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("imrcd_gltf_load", C.c_int, [_P, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]),
    ("imrcd_mesh_begin", C.c_int, [_P]),
    ("imrcd_mesh_add_primitive", C.c_int, [_P, _P, C.c_uint64, C.c_uint32, _P, _P, C.c_uint64, C.c_uint32]),
    ("imrcd_mesh_end", C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("imrcd_mesh_import_tree", C.c_int, [_P, C.c_uint64, _P, _P, _P, _P, _P, C.c_uint64, _P, _P, _P, _P, C.POINTER(C.c_uint32)]),
    ("imrcd_mesh_info", C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("imrcd_mesh_export_tree", C.c_int, [_P, C.c_uint32] + [_P] * 9),
    ("imrcd_mesh_last_build_ms", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("imrcd_mesh_update_positions", C.c_int, [_P, C.c_uint32, _P, _P]),
    ("imrcd_mesh_refit", C.c_int, [_P, _P, C.c_uint64]),
    ("imrcd_mesh_last_refit_ms", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("imrcd_skin_create", C.c_int, [_P, C.c_uint64, C.c_uint32, _P, C.c_uint32, _P, _P, C.POINTER(C.c_uint32)]),
    ("imrcd_mesh_bind_skin", C.c_int, [_P, C.c_uint32, C.c_uint32]),
    ("imrcd_meshes_repose", C.c_int, [_P, C.c_uint64, _P, _P, _P, _P, _P]),
    ("imrcd_mesh_last_repose_ms", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("imrcd_test_reposed_vertices", C.c_int, [_P, _P, C.c_uint64]),
    ("imrcd_frame_reset", C.c_int, [_P]),
    ("imrcd_frame_add_entry", C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint8, C.c_uint32]),
    ("imrcd_frame_add_entries", C.c_int, [_P, C.c_uint64, _P, _P, _P, _P, _P]),
    ("imrcd_frame_map_entries", C.c_int, [_P, C.c_uint64, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    ("imrcd_frame_commit_entries", C.c_int, [_P, C.c_uint64, C.c_int]),
    ("imrcd_frame_set_shard", C.c_int, [_P, C.c_uint32, C.c_uint32]),
    ("imrcd_frame_execute", C.c_int, [_P]),
    ("imrcd_frame_upload", C.c_int, [_P]),
    ("imrcd_frame_run", C.c_int, [_P]),
    ("imrcd_frame_fetch", C.c_int, [_P]),
    ("imrcd_frame_run_async", C.c_int, [_P]),
    ("imrcd_frame_finish", C.c_int, [_P]),
    ("imrcd_frame_results", C.c_int, [_P, C.POINTER(C.POINTER(EntityPair)), C.POINTER(C.c_uint64), C.POINTER(C.POINTER(TriHit)), C.POINTER(C.c_uint64)]),
    ("imrcd_frame_pairs", C.c_int, [_P, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]),
    ("imrcd_frame_combos", C.c_int, [_P, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]),
    ("imrcd_frame_get_stats", C.c_int, [_P, C.POINTER(FrameStats)]),
    ("imrcd_frame_results_device", C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64), C.POINTER(_P), C.POINTER(C.c_uint64)]),
    ("imrcd_frame_results_block", C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("imrcd_frame_results_local", C.c_int, [_P, C.POINTER(C.POINTER(EntityPair)), C.POINTER(C.c_uint64)]),
    ("imrcd_comm_unique_id", C.c_int, [_P]),
    ("imrcd_comm_init", C.c_int, [_P, _P, C.c_uint32, C.c_uint32]),
    ("imrcd_comm_destroy", C.c_int, [_P]),
    ("imrcd_comm_transport", C.c_int, [_P]),
    ("imrcd_group_create", C.c_int, [C.POINTER(C.c_int), C.c_uint32, C.POINTER(_P)]),
    ("imrcd_group_destroy", None, [_P]),
    ("imrcd_group_size", C.c_uint32, [_P]),
    ("imrcd_group_ctx", _P, [_P, C.c_uint32]),
    ("imrcd_group_last_error", C.c_char_p, [_P]),
    ("imrcd_group_mesh_create", C.c_int, [_P, _P, _P, _P, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("imrcd_group_gltf_load", C.c_int, [_P, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]),
    ("imrcd_group_frame_reset", C.c_int, [_P]),
    ("imrcd_group_frame_add_entry", C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint8, C.c_uint32]),
    ("imrcd_group_frame_add_entries", C.c_int, [_P, C.c_uint64, _P, _P, _P, _P, _P]),
    ("imrcd_group_frame_execute", C.c_int, [_P]),
    ("imrcd_group_frame_results", C.c_int, [_P, C.POINTER(C.POINTER(EntityPair)), C.POINTER(C.c_uint64)]),
    ("imrcd_test_sat", C.c_int, [_P, C.c_uint64, _P, _P, _P, _P, _P, _P]),
    ("imrcd_test_tri_tri", C.c_int, [_P, C.c_uint64, _P, _P, _P, _P, _P]),
    ("imrcd_test_pair_matrix", C.c_int, [_P, C.c_uint64, _P, _P, _P]),
    ("imrcd_test_obb_fit", C.c_int, [_P, C.c_uint64, _P, _P]),
    ("imrcd_test_ray_tree", C.c_int, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P, _P, _P]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGS]

_lib = None


def preload_nccl():
    """The library binds NCCL at run time by soname (dlopen "libnccl.so.2").  A process that will also import torch must end up with ONE
    NCCL: load torch's bundled copy first (when there is one) so that the soname resolves to it, whichever of the two is used first."""
    import glob
    import site
    for sp in site.getsitepackages():
        for path in glob.glob(os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2")):
            try:
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return path
            except OSError:
                pass
    return None


def load() -> C.CDLL:
    """dlopen libimrcd.so and bind every declared symbol.  Never touches the GPU by itself."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or `make -C inmyroom_vulkan_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(SO_PATH)
    for name, res, args in _SIGS:
        fn = getattr(lib, name)       # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
