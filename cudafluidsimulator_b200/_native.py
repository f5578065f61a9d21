"""ctypes binding of libsph_b200.so (the C ABI declared in include/sph_b200.h).

There is no Python or CPU fallback: if the shared library is missing the import
fails loudly, and every call that needs a GPU returns the CUDA error.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# SPH_B200_LIB selects another build of the same ABI (e.g. the self-checking one)
LIB_PATH = Path(os.environ.get("SPH_B200_LIB", PKG / "libsph_b200.so"))

SPH_KEY_FLAT = 0
SPH_KEY_MORTON = 1
SPH_SORT_COUNT = 1
SPH_SORT_RADIX = 2
SPH_STAGE_COUNT = 8


class SphSettings(C.Structure):
    """Layout-identical to the reference `struct Settings` (ref: simulator.h:19-31)."""
    _fields_ = [
        ("randomInit", C.c_uint8), ("_pad", C.c_uint8 * 3),
        ("numParticles", C.c_int32),
        ("h", C.c_float),
        ("v_kernel_coeff", C.c_float),
        ("d_kernel_coeff", C.c_float),
        ("boxDim", C.c_float),
        ("numCellsPerDim", C.c_float),
        ("timestep", C.c_float),
    ]


class SphTimes(C.Structure):
    """Layout-identical to the reference `struct Times` (ref: times.h:5-10)."""
    _fields_ = [("buildGrid", C.c_double), ("sphUpdate", C.c_double), ("memcpy", C.c_double),
                ("iters", C.c_int32)]


class SphOptions(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("key_mode", C.c_int32), ("record_force", C.c_int32),
        ("use_graph", C.c_int32), ("capacity", C.c_int32),
        ("z_cell_lo", C.c_int32), ("z_cell_hi", C.c_int32),
        ("no_mask_handoff", C.c_int32),
        ("nz_cells", C.c_int32), ("ghost_capacity", C.c_int32), ("emig_capacity", C.c_int32),
        ("pipeline_readback", C.c_int32), ("stage_tiles", C.c_int32),
        ("density_sum", C.c_int32), ("sort_algo", C.c_int32), ("reserved", C.c_int32 * 1),
    ]


SPH_NCCL_ID_BYTES = 128
SPH_MAX_LOCAL_SLABS = 16


class SphClusterOptions(C.Structure):
    _fields_ = [("world", C.c_int32), ("first_rank", C.c_int32), ("local_count", C.c_int32),
                ("devices", C.c_int32 * SPH_MAX_LOCAL_SLABS), ("nz_cells", C.c_int32),
                ("capacity", C.c_int32), ("ghost_capacity", C.c_int32), ("emig_capacity", C.c_int32),
                ("density_sum", C.c_int32), ("rebalance_every", C.c_int32), ("reserved", C.c_int32 * 6),
                ("nccl_id", C.c_uint8 * SPH_NCCL_ID_BYTES)]


class SphSlabStats(C.Structure):
    _fields_ = [("rank", C.c_int32), ("device", C.c_int32), ("z_cell_lo", C.c_int32), ("z_cell_hi", C.c_int32),
                ("n_owned", C.c_int32), ("ghosts_lo", C.c_int32), ("ghosts_hi", C.c_int32), ("steps", C.c_int32),
                ("migrated_total", C.c_int64), ("ghosts_total", C.c_int64), ("overflow", C.c_uint32),
                ("rebalances", C.c_int32), ("kinetic_energy", C.c_double), ("density_sum", C.c_double),
                ("debug_flags", C.c_uint32), ("checked_build", C.c_int32)]


# every symbol include/sph_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F = C.POINTER(C.c_float)
_U = C.POINTER(C.c_uint32)
_I = C.POINTER(C.c_int32)
_D = C.POINTER(C.c_double)
SYMBOLS = {
    "sph_create": (C.c_int, [C.POINTER(SphSettings), C.POINTER(_P)]),
    "sph_create_ex": (C.c_int, [C.POINTER(SphSettings), C.POINTER(SphOptions), C.POINTER(_P)]),
    "sph_destroy": (None, [_P]),
    "sph_setup": (C.c_int, [_P]),
    "sph_step": (C.c_int, [_P]),
    "sph_step_timed": (C.c_int, [_P, C.POINTER(SphTimes)]),
    "sph_advance": (C.c_int, [_P, C.c_int]),
    "sph_advance_timed": (C.c_int, [_P, C.c_int, _F]),
    "sph_push": (C.c_int, [_P, C.c_int, C.c_int]),
    "sph_positions_host": (_F, [_P]),
    "sph_readback": (C.c_int, [_P]),
    "sph_set_state": (C.c_int, [_P, _F, _F]),
    "sph_get_state": (C.c_int, [_P, _F, _F]),
    "sph_get_keys": (C.c_int, [_P, C.c_int, _U]),
    "sph_get_sorted_index": (C.c_int, [_P, _U, _U]),
    "sph_get_cell_start": (C.c_int, [_P, _U, _U]),
    "sph_get_neighbor_counts": (C.c_int, [_P, _I, _I]),
    "sph_get_density_pressure_force": (C.c_int, [_P, _F, _F, _F]),
    "sph_get_stats": (C.c_int, [_P, _D, _D]),
    "sph_slab_load": (C.c_int, [_P, C.c_int, _F, _F, _U]),
    "sph_cluster_nccl_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "sph_cluster_create": (C.c_int, [C.POINTER(SphSettings), C.POINTER(SphClusterOptions), C.POINTER(_P)]),
    "sph_cluster_destroy": (None, [_P]),
    "sph_cluster_setup": (C.c_int, [_P]),
    "sph_cluster_load": (C.c_int, [_P, C.c_int, C.c_int, _F, _F, _U]),
    "sph_cluster_advance": (C.c_int, [_P, C.c_int]),
    "sph_cluster_advance_timed": (C.c_int, [_P, C.c_int, _F]),
    "sph_cluster_step": (C.c_int, [_P]),
    "sph_cluster_sync": (C.c_int, [_P]),
    "sph_cluster_host_records": (C.c_int, [_P, C.c_int, C.POINTER(_F), _I]),
    "sph_cluster_positions": (C.c_int, [_P, _F, C.c_int64]),
    "sph_cluster_download": (C.c_int, [_P, C.c_int, _U, _F, _F, _I]),
    "sph_cluster_stats": (C.c_int, [_P, C.c_int, C.POINTER(SphSlabStats)]),
    "sph_cluster_rebalance": (C.c_int, [_P]),
    "sph_cluster_launch_count": (C.c_int64, [_P]),
    "sph_debug_flags": (C.c_int, [_P, _U, _I]),
    "sph_profile_enable": (C.c_int, [_P, C.c_int]),
    "sph_profile_read": (C.c_int, [_P, _D, C.POINTER(C.c_int64), C.c_int]),
    "sph_stage_name": (C.c_char_p, [C.c_int]),
    "sph_launch_count": (C.c_int64, [_P]),
    "sph_num_particles": (C.c_int, [_P]),
    "sph_sort_info": (C.c_int, [_P, _I, _I, _I, _I]),
    "sph_last_error": (C.c_char_p, []),
    "sph_abi_version": (C.c_int, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libsph_b200.so (built in-tree by cudafluidsimulator_b200.build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m cudafluidsimulator_b200.build` "
                "(there is no CPU fallback for the SPH step)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class SphError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"sph_b200 error {code}: {message}")
        self.code = code


def check(code: int) -> None:
    if code != 0:
        raise SphError(code, load().sph_last_error().decode(errors="replace"))
