"""ctypes loader for libseqwin_b200.so (the C ABI declared in include/seqwin_b200.h)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libseqwin_b200.so"

SW_OK, SW_ERR_RUNTIME, SW_ERR_VALUE = 0, 1, 2
SW_KMERS, SW_NODES, SW_EDGES, SW_OFFSETS, SW_RECORDS = range(5)


class StageTimes(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "sketch_ms", "sort_nodes_ms", "nodes_ms", "edges_ms",
                                         "d2h_ms", "total_ms", "plan_ms", "sketch_kernel_ms", "reorder_ms", "penalty_ms")] + \
               [(n, C.c_uint64) for n in ("n_bases", "n_kmers", "n_nodes", "n_edges", "n_tiles",
                                          "sketch_launches", "total_launches")]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/seqwin_b200.h declares: (restype, argtypes)
_P, _SZ, _U32, _I = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
NODES_READY_FN = C.CFUNCTYPE(None, _P, _P)   # sw_nodes_ready_fn(user, graph)
PROTOTYPES = {
    "sw_last_error": (C.c_char_p, []),
    "sw_build": (_I, [C.POINTER(C.c_char_p), _SZ, _U32, _U32, _U32, _I, C.POINTER(_P)]),
    "sw_graph_size": (_SZ, [_P, _I]),
    "sw_graph_export": (_I, [_P, _P, _P, _P, _P]),
    "sw_graph_n_records": (_SZ, [_P, _SZ]),
    "sw_graph_record_id": (C.c_char_p, [_P, _SZ, _SZ]),
    "sw_graph_free": (None, [_P]),
    "sw_get_penalty": (_I, [_P, _SZ, _P, _SZ, _P, _SZ, _P, _SZ, _U32]),
    "sw_filter_kmers": (_I, [_P, _SZ, _P, _SZ, _P, _SZ, _P, _P, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "sw_batch_from_fasta": (_I, [C.POINTER(C.c_char_p), _SZ, _U32, C.POINTER(_P)]),
    "sw_batch_from_memory": (_I, [C.POINTER(_P), _P, _P, C.POINTER(C.c_char_p), _SZ, _SZ, _U32, C.POINTER(_P)]),
    "sw_batch_concat": (_I, [C.POINTER(_P), _SZ, C.POINTER(_P)]),
    "sw_batch_n_bases": (_SZ, [_P]),
    "sw_batch_n_records": (_SZ, [_P]),
    "sw_batch_packed_bytes": (_SZ, [_P]),
    "sw_batch_record_offsets": (_I, [_P, _P, _SZ]),
    "sw_batch_free": (None, [_P]),
    "sw_dev_upload": (_I, [_P, C.POINTER(_P)]),
    "sw_dev_batch_free": (None, [_P]),
    "sw_dev_build": (_I, [_P, _U32, _U32, C.POINTER(_P), C.POINTER(StageTimes)]),
    "sw_build_from_batch": (_I, [_P, _U32, _U32, C.POINTER(_P), C.POINTER(StageTimes)]),
    "sw_build_from_batch_scored": (_I, [_P, _U32, _U32, _P, _SZ, C.POINTER(_P), C.POINTER(StageTimes)]),
    "sw_dev_build_scored": (_I, [_P, _U32, _U32, _P, _SZ, C.POINTER(_P), C.POINTER(StageTimes)]),
    "sw_graph_penalty": (_I, [_P, _P, _SZ, _P, _SZ, C.POINTER(C.c_float)]),
    "sw_dev_sketch": (_I, [_P, _U32, _U32, _P, _P, _P, _SZ, C.POINTER(_SZ)]),
    "sw_set_stream": (_I, [_P]),
    "sw_dev_build_ex": (_I, [_P, _U32, _U32, _U32, _P, _SZ, C.POINTER(_P), C.POINTER(StageTimes)]),
    "sw_build_from_batch_ex": (_I, [_P, _U32, _U32, _U32, _I, _P, _SZ, C.POINTER(_P), C.POINTER(StageTimes)]),
    "sw_graph_finish_penalty": (_I, [_P, C.c_uint64, C.c_uint64]),
    "sw_graph_fetch": (_I, [_P]),
    "sw_graph_device_ptrs": (_I, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "sw_graph_split": (_I, [_P, _U32, _P, _P, _P]),
    "sw_set_nodes_ready": (_I, [NODES_READY_FN, _P]),
    "sw_dist_merge": (_I, [_P, _P, _P, _P, _P, _P, _P, _U32, C.POINTER(_P), C.POINTER(_U32)]),
    "sw_graph_filter_edges": (_I, [_P, C.c_uint64]),
    "sw_filter_edges_and_nodes": (_I, [_P, _SZ, _P, _SZ, C.c_uint64, _P, _P, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "sw_graph_filter_kmers": (_I, [_P, _P, _SZ]),
    "sw_dist_merge_edges": (_I, [_P, _P, _P, _U32, C.POINTER(_U32)]),
    "sw_graph_count_sums": (_I, [_P, _P]),
    "sw_mem_stats": (_I, [C.POINTER(C.c_uint64)]),
    "sw_dev_sketch_route": (_I, [_P, _U32, _U32, _U32, C.POINTER(_P), C.POINTER(StageTimes)]),
    "sw_routed_info": (_I, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_uint64), _P, _P]),
    "sw_routed_free": (None, [_P]),
    "sw_dev_aggregate": (_I, [_P, _P, _P, _P, C.c_uint64, _U32, _U32, _P, _SZ, _P, _SZ, C.c_double, _P, _U32, C.POINTER(_P),
                              C.POINTER(StageTimes)]),
    "sw_peer_alloc": (_I, [_SZ, C.POINTER(_P), _P]),
    "sw_peer_open": (_I, [_P, C.POINTER(_P)]),
    "sw_peer_close": (_I, [_P]),
    "sw_peer_free": (_I, [_P]),
    "sw_dev_sketch_hist": (_I, [_P, _U32, _U32, _U32, _U32, C.POINTER(_P), _P, C.POINTER(StageTimes)]),
    "sw_routed_scatter": (_I, [_P, _P, _P, C.POINTER(C.c_float)]),
    "sw_set_low_memory": (_I, [_I]),
    "sw_trim_memory": (_I, []),
    "sw_measure_int_peak": (_I, [C.POINTER(C.c_double)]),
    "sw_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_SZ)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; fail loudly if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m seqwin_b200.build` "
                "(nvcc, sm_100a). seqwin_b200 has no CPU fallback.")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    """Map C-ABI return codes onto the exception types the reference raises."""
    if rc == SW_OK:
        return
    msg = (lib().sw_last_error() or b"").decode(errors="replace")
    if rc == SW_ERR_VALUE:
        raise ValueError(msg)
    raise RuntimeError(msg)
