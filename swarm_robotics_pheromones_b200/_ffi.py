"""ctypes binding of libpherofield.so (include/pherofield.h). No CPU fallback: if the library is
missing or cannot be loaded, importing the binding raises."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

from . import abi, build

_LIB = None

_vp = C.c_void_p
_cfgp = C.POINTER(abi.PFConfig)
_agp = C.POINTER(abi.PFAgents)

# name -> (restype, argtypes); must list every symbol include/pherofield.h declares
SIGNATURES = {
    "pf_abi_version": (C.c_int, []),
    "pf_config_default": (C.c_int, [_cfgp, C.c_int32, C.c_int32, C.c_int32]),
    "pf_create": (C.c_int, [_cfgp, C.POINTER(_vp)]),
    "pf_destroy": (C.c_int, [_vp]),
    "pf_last_error": (C.c_char_p, [_vp]),
    "pf_get_config": (C.c_int, [_vp, _cfgp]),
    "pf_reserve_agents": (C.c_int, [_vp, C.c_int64]),
    "pf_field_ptr": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "pf_field_fill": (C.c_int, [_vp, C.c_int, C.c_float, _vp]),
    "pf_field_upload_host": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "pf_snapshot_host": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "pf_fill_ghosts": (C.c_int, [_vp, _vp]),
    "pf_field_sum_host": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), _vp]),
    "pf_deposit": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "pf_agents_deposit": (C.c_int, [_vp, _agp, _vp]),
    "pf_step_field": (C.c_int, [_vp, C.c_int, _vp]),
    "pf_step_field_rows": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "pf_swap": (C.c_int, [_vp]),
    "pf_sense": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pf_agents_step": (C.c_int, [_vp, _agp, C.c_uint32, _vp, _vp]),
    "pf_step": (C.c_int, [_vp, _agp, C.c_int, _vp]),
    "pf_step_ex": (C.c_int, [_vp, _agp, C.c_int, C.c_uint32, _vp]),
    "pf_agents_observe": (C.c_int, [_vp, _agp, _vp, _vp, _vp, _vp]),
    "pf_agents_sort": (C.c_int, [_vp, _agp, _vp]),
    "pf_set_time": (C.c_int, [_vp, C.c_double, C.c_double]),
    "pf_get_time": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "pf_evaporate_trail": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_double, _vp, _vp]),
    "pf_migrate_pack": (C.c_int, [_vp, _agp, _vp, _vp, C.c_int64, _vp]),
    "pf_migrate_unpack": (C.c_int, [_vp, _agp, _vp, C.c_int64, _vp]),
    "pf_migrate_unpack2": (C.c_int, [_vp, _agp, _vp, _vp, C.c_int64, _vp]),
    "pf_peer_export": (C.c_int, [_vp, C.POINTER(abi.PFPeerInfo)]),
    "pf_peer_ipc_handles": (C.c_int, [_vp, _vp]),
    "pf_peer_ipc_open": (C.c_int, [_vp, _vp, C.POINTER(abi.PFPeerInfo), C.POINTER(abi.PFPeerInfo)]),
    "pf_peer_attach": (C.c_int, [_vp, C.POINTER(abi.PFPeerInfo), C.POINTER(abi.PFPeerInfo)]),
    "pf_peer_push_edges": (C.c_int, [_vp, _vp]),
    "pf_peer_signal": (C.c_int, [_vp, _vp]),
    "pf_peer_wait": (C.c_int, [_vp, _vp]),
    "pf_peer_status": (C.c_int, [_vp, _vp]),
    "pf_strip_tick": (C.c_int, [_vp, _agp, C.c_uint32, C.c_uint32, _vp]),
    "pf_strip_profile": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "pf_counters_host": (C.c_int, [_vp, C.POINTER(abi.PFCounters), _vp]),
    "pf_counters_reset": (C.c_int, [_vp, _vp]),
    "pf_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "pf_tile_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pf_tile_dense": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pf_tile_rows": (C.c_int, [_vp, _agp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pf_profile_enable": (C.c_int, [_vp, C.c_int]),
    "pf_profile_read": (C.c_int, [_vp, C.POINTER(abi.PFProfile), C.c_int]),
    "pf_trail_create": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.POINTER(_vp)]),
    "pf_trail_destroy": (C.c_int, [_vp]),
    "pf_trail_last_error": (C.c_char_p, [_vp]),
    "pf_trail_configure": (C.c_int, [_vp, _vp, _vp, _vp]),
    "pf_trail_set_state": (C.c_int, [_vp, C.c_int64, C.c_int]),
    "pf_trail_tick": (C.c_int, [_vp, C.c_double, C.c_double, _vp, _vp, C.POINTER(abi.PFTrailOut), _vp]),
    "pf_trail_dropped": (C.c_int, [_vp, C.POINTER(C.c_uint64), _vp]),
    "pf_trail_entries": (C.c_int, [_vp] + [C.POINTER(_vp)] * 6 + [C.POINTER(C.c_int)]),
    "pf_heatmap_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    "pf_heatmap_destroy": (C.c_int, [_vp]),
    "pf_heatmap_last_error": (C.c_char_p, [_vp]),
    "pf_heatmap_accumulate": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_double, _vp]),
    "pf_heatmap_snapshot_host": (C.c_int, [_vp, _vp, _vp]),
    "pf_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "pf_set_stencil_variant": (C.c_int, [_vp, C.c_int, C.c_int]),
}


def header_symbols() -> list[str]:
    """every function name declared in include/pherofield.h"""
    text = build.HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", text)))


def library_path() -> Path:
    return build.LIB


def lib() -> C.CDLL:
    """Load (building first when stale and nvcc is present) the C-ABI library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = build.LIB
    if build.is_stale():
        try:
            build.build()
        except Exception as e:  # no nvcc on this box: use the shipped .so if there is one
            if not path.exists():
                raise RuntimeError(
                    f"libpherofield.so is not built and cannot be built here ({e}); "
                    "pherofield has no CPU fallback") from e
            import warnings
            warnings.warn(f"libpherofield.so is older than its sources and could not be rebuilt ({e}); "
                          "loading the existing library", RuntimeWarning)
    L = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError = symbol missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if L.pf_abi_version() != abi.PF_ABI_VERSION:
        raise RuntimeError("libpherofield.so ABI version mismatch; rebuild")
    _LIB = L
    return L


class PFError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pherofield error {code}: {msg}")
        self.code = code


def check(rc: int, handle=None):
    if rc != 0:
        msg = lib().pf_last_error(handle)
        raise PFError(rc, msg.decode() if msg else "")
