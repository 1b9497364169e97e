"""ctypes loader of libbbcp.so (the C ABI of include/bbcp.h). Fails loudly: there is no fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BBCP_LIB") or os.path.join(_HERE, "libbbcp.so")   # BBCP_LIB: a diagnostic build (make checks)


class Opts(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("small_trail", C.c_uint32), ("out_capacity", C.c_uint64), ("mid_trail", C.c_uint32), ("merge_words_per_rank", C.c_uint32)]


class BatchInfo(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_implied", C.c_uint64), ("n_ok", C.c_uint64), ("n_conflict", C.c_uint64),
                ("n_skipped", C.c_uint64), ("n_large", C.c_uint64), ("props_all", C.c_uint64), ("props_ref", C.c_uint64),
                ("slots", C.c_uint64), ("clause_fetches", C.c_uint64), ("kernel_bytes", C.c_uint64),
                ("triage_bytes", C.c_uint64), ("n_deep", C.c_uint64),
                ("kernel_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_uint32), ("tier1_ms", C.c_float),
                ("triage_ms", C.c_float), ("deep_ms", C.c_float), ("block_ms", C.c_float), ("mid_ms", C.c_float), ("n_mid", C.c_uint64),
                ("compact_ms", C.c_float), ("merge_ms", C.c_float),
                ("adopted", C.c_uint64), ("adopt_lits", C.c_uint64), ("inherited", C.c_uint64), ("assignments", C.c_uint64), ("bcps", C.c_uint64),
                ("assignments_ok", C.c_uint64), ("props_ref_ok", C.c_uint64), ("hot_rows", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class DbInfo(C.Structure):
    _fields_ = [("max_var", C.c_int32), ("n_mapped_vars", C.c_uint64), ("n_level0", C.c_uint64), ("n_bin", C.c_uint64),
                ("n_tern", C.c_uint64), ("n_long", C.c_uint64), ("n_long_lits", C.c_uint64), ("row_slots", C.c_uint64),
                ("arena_words", C.c_uint64), ("device_bytes", C.c_uint64), ("db_version", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_SIGS = {
    "bbcp_create": (C.c_void_p, [C.c_int32]),
    "bbcp_release": (None, [C.c_void_p]),
    "bbcp_set_device": (C.c_int, [C.c_void_p, C.c_int]),
    "bbcp_add": (C.c_int, [C.c_void_p, C.c_int32]),
    "bbcp_add_clause": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "bbcp_add_clauses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "bbcp_finalize": (C.c_int, [C.c_void_p]),
    "bbcp_host_prepare": (C.c_int, [C.c_void_p]),
    "bbcp_get_db_info": (C.c_int, [C.c_void_p, C.POINTER(DbInfo)]),
    "bbcp_level0_lits": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "bbcp_unit_log": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "bbcp_host_row": (C.c_uint64, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]),
    "bbcp_host_order": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "bbcp_status": (C.c_int, [C.c_void_p]),
    "bbcp_error_string": (C.c_char_p, [C.c_void_p]),
    "bbcp_propagate_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Opts), C.POINTER(BatchInfo)]),
    "bbcp_probe_lits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Opts), C.POINTER(BatchInfo)]),
    "bbcp_probe_lits_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Opts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.POINTER(BatchInfo)]),
    "bbcp_probe_range": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(Opts), C.POINTER(BatchInfo)]),
    "bbcp_probe_all": (C.c_int, [C.c_void_p, C.POINTER(Opts), C.POINTER(BatchInfo)]),
    "bbcp_propagate_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(Opts), C.POINTER(BatchInfo)]),
    "bbcp_probe_to_fixpoint": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "bbcp_probe_products": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "bbcp_check_fixpoints": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bbcp_probe_hbr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "bbcp_fetch_flags": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bbcp_fetch_implied": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bbcp_fetch_implied32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bbcp_bitmap_words": (C.c_uint64, [C.c_void_p]),
    "bbcp_fetch_bitmaps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bbcp_clear_bitmaps": (C.c_int, [C.c_void_p]),
    "bbcp_device_failed_bitmap": (C.c_void_p, [C.c_void_p]),
    "bbcp_device_unit_bitmap": (C.c_void_p, [C.c_void_p]),
    "bbcp_device_flags": (C.c_void_p, [C.c_void_p]),
    "bbcp_device_impl_off": (C.c_void_p, [C.c_void_p]),
    "bbcp_device_impl_lits": (C.c_void_p, [C.c_void_p]),
    "bbcp_stream": (C.c_void_p, [C.c_void_p]),
    "bbcp_sync": (C.c_int, [C.c_void_p]),
    "bbcp_debug_phase_ticks": (None, [C.c_void_p, C.c_void_p]),
    "bbcp_l2_flush": (C.c_int, [C.c_void_p, C.c_uint64]),
    "bbcp_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bbcp_peer_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "bbcp_peer_detach": (C.c_int, [C.c_void_p]),
    "bbcp_merge_bitmaps": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_float)]),
    "bbcp_assume": (C.c_int, [C.c_void_p, C.c_int32]),
    "bbcp_solve": (C.c_int, [C.c_void_p]),
    "bbcp_val": (C.c_int, [C.c_void_p, C.c_int32]),
    "bbcp_declevel": (C.c_int, [C.c_void_p, C.c_int32]),
    "bbcp_backtrack": (C.c_int, [C.c_void_p, C.c_int32]),
    "bbcp_propagations": (C.c_uint64, [C.c_void_p]),
}

PEER_HANDLE_BYTES = 128
_LIB = None


def load() -> C.CDLL:
    """Load the CUDA engine. Raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``
    or ``make -C intel_sat_solver_b200/csrc``): there is deliberately no CPU or pure-Python substitute."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA extension first (make -C intel_sat_solver_b200/csrc); "
                               "this engine has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _LIB = lib
    return _LIB


def exported_names():
    return sorted(_SIGS)
