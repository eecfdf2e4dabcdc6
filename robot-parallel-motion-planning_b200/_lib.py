"""ctypes binding of libgmt_b200.so -- one prototype per entry point of include/gmt_c_api.h."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMT_B200_LIB") or os.path.join(HERE, "libgmt_b200.so")   # override: tuning builds only

GMT_OK = 0
GMT_STATUS_GOAL = 0
GMT_STATUS_LIMIT = 1


class GmtResult(C.Structure):
    """struct gmt_result (include/gmt_c_api.h)."""
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("goal_cost", C.c_float),
                ("route_len", C.c_int), ("device_ms", C.c_float), ("grid_overflow", C.c_int),
                ("edge_pairs", C.c_longlong), ("edge_evals", C.c_longlong),
                ("word_checks", C.c_longlong), ("launches", C.c_int), ("farthest_frontier", C.c_int)]


class GmtError(RuntimeError):
    pass


# every symbol include/gmt_c_api.h declares: name -> (restype, argtypes)
_vp, _i, _f, _ip, _fp = C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_float)
PROTOTYPES = {
    "gmt_create": (_i, [_vp, _i, _vp, _vp, _i, C.POINTER(_vp)]),
    "gmt_destroy": (_i, [_vp]),
    "gmt_set_stream": (_i, [_vp, _vp]),
    "gmt_plan": (_i, [_vp, _i, _i, _f, _f, _vp, _i, _i, C.POINTER(GmtResult)]),
    "gmt_set_obstacles": (_i, [_vp, _vp, _i]),
    "gmt_plan_resident": (_i, [_vp, _i, _i, _f, _f, _i, C.POINTER(GmtResult)]),
    "gmt_update_obstacles": (_i, [_vp, _vp, _vp, _i]),
    "gmt_set_words": (_i, [_vp, _i]),
    "gmt_set_obstacles_oriented": (_i, [_vp, _vp, _i]),
    "gmt_get_route": (_i, [_vp, _vp, _i, _ip]),
    "gmt_get_route_edges": (_i, [_vp, _vp, _vp, _vp, _i, _ip]),
    "gmt_get_parent": (_i, [_vp, _vp]),
    "gmt_get_cost": (_i, [_vp, _vp]),
    "gmt_set_trace": (_i, [_vp, _i]),
    "gmt_get_trace": (_i, [_vp, _vp, _i, _vp, C.c_longlong, C.POINTER(C.c_longlong), _vp]),
    "gmt_get_phase_times": (_i, [_vp, _vp, _i]),
    "gmt_edge_costs": (_i, [_vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _fp]),
    "gmt_sample_states": (_i, [C.c_uint, _i, _f, _i, _i, _vp]),
    "gmt_build_neighbors": (_i, [_vp, _i, C.c_double, _i, C.POINTER(_vp), C.POINTER(_vp),
                                 C.POINTER(C.c_longlong)]),
    "gmt_free": (None, [_vp]),
    "gmt_insert_states": (_i, [_vp, _vp, _i, C.c_double, _ip]),
    "gmt_reset_inserted": (_i, [_vp]),
    "gmt_plan_batch": (_i, [_vp, _vp, _vp, _i, _f, _f, _i, _vp, _vp, _i]),
    "gmt_sharded_begin": (_i, [_vp, _i, _i, _f, _f, _i, _i, _i]),
    "gmt_sharded_step": (_i, [_vp, _ip]),
    "gmt_best_dev": (_i, [_vp, C.POINTER(_vp), C.POINTER(C.c_longlong)]),
    "gmt_sharded_finish": (_i, [_vp, C.POINTER(GmtResult)]),
    "gmt_peer_export": (_i, [_vp, _vp]),
    "gmt_peer_open": (_i, [_vp, _i, _i, _vp]),
    "gmt_peer_close": (_i, [_vp]),
    "gmt_plan_sharded_p2p": (_i, [_vp, _i, _i, _f, _f, _i, _i, _i, C.POINTER(GmtResult)]),
    "gmt_create_from_states": (_i, [_vp, _i, C.c_double, _i, C.POINTER(_vp)]),
    "gmt_get_neighbors": (_i, [_vp, _vp, _vp, C.c_longlong, _vp]),
    "gmt_set_obstacles_from_vehicles": (_i, [_vp, _vp, _i, C.c_double, C.c_double, C.c_double, C.c_double, _i, _vp]),
    "gmt_get_obstacles": (_i, [_vp, _vp, _vp, _i, _vp]),
    "gmt_set_queries": (_i, [_vp, _vp, _vp, _i]),
    "gmt_plan_queries": (_i, [_vp, _f, _f, _i, _vp, _vp, _i]),
    "gmt_debug_batch_times": (_i, [_vp, _vp, _i, _vp]),
    "gmt_set_option": (_i, [_vp, C.c_char_p, _i]),
    "gmt_debug_barrier_us": (_i, [_vp, _i, _i, _i, _fp]),
    "gmt_debug_counters": (_i, [_vp, _vp]),
    "gmt_debug_sincos_check": (_i, [_i, _vp, _vp]),
    "gmt_measure_fp32_peak": (_i, [_i, _fp]),
    "gmt_last_error": (C.c_char_p, []),
    "gmt_version": (_i, []),
}

_lib = None


def load():
    """Load the CUDA library; there is deliberately no fallback when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GmtError("%s is missing -- build it with `python %s` (nvcc, sm_100a); this planner "
                           "has no CPU fallback" % (LIB_PATH, os.path.join(HERE, "build.py")))
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != GMT_OK:
        msg = load().gmt_last_error()
        raise GmtError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
