"""ctypes binding of libs1s2_b200.so (the C ABI in include/s1s2_b200.h).

There is no fallback: if the shared library is missing or does not load, importing the compute entry points raises.
"""
import ctypes as C
import os

from ._build import LIB_PATH

OK, ERR_INVALID, ERR_CUDA, ERR_STATE = 0, 1, 2, 3
STEP_NONE, STEP_EPS_DDIM, STEP_V_DDIM, STEP_EPS_DDPM, STEP_V_DDPM = 0, 1, 2, 3, 4
STEP_FINAL, STEP_NOISE, STEP_PHILOX = 1, 2, 4


class Step(C.Structure):
    """struct s1s2_step"""
    _fields_ = [("t", C.c_int32), ("kind", C.c_int32), ("flags", C.c_int32), ("noise_index", C.c_int32),
                ("c0", C.c_float), ("c1", C.c_float), ("c2", C.c_float), ("c3", C.c_float), ("c4", C.c_float)]


class S1S2Error(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); mirrors include/s1s2_b200.h one to one
SIGNATURES = {
    "s1s2_abi_version": (C.c_int, []),
    "s1s2_global_error": (C.c_char_p, []),
    "s1s2_last_error": (C.c_char_p, [C.c_void_p]),
    "s1s2_launch_count": (C.c_int64, [C.c_void_p]),
    "s1s2_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "s1s2_destroy": (None, [C.c_void_p]),
    "s1s2_load_weights": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int64), C.c_void_p]),
    "s1s2_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "s1s2_sample": (C.c_int, [C.c_void_p, C.POINTER(Step), C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "s1s2_set_noise_seed": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32]),
    "s1s2_sample_host": (C.c_int, [C.c_void_p, C.POINTER(Step), C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_int, C.c_void_p]),
    "s1s2_sample_host_stream": (C.c_int, [C.c_void_p, C.POINTER(Step), C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                          C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "s1s2_patch_noise": (C.c_int, [C.c_int, C.c_uint64, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "s1s2_debug_activation": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "s1s2_profile_layers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int),
                                      C.c_void_p]),
    "s1s2_debug_loop_layer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "s1s2_debug_saturation_count": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int),
                                              C.c_void_p]),
    "s1s2_view_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "s1s2_debug_tile_width": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "s1s2_layer_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "s1s2_tile_extract": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "s1s2_tile_filter": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_int, C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "s1s2_patch_metrics": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "s1s2_stitch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "s1s2_stitch_weighted": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def lib():
    """The loaded library (loads on first use; raises if it is not built)."""
    global _lib
    if _lib is None:
        path = os.environ.get("S1S2_LIB", LIB_PATH)      # override: same-box A/B of two builds (measurement aid)
        if not os.path.exists(path):
            raise S1S2Error(f"{path} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "s1s2_b200 has no CPU or PyTorch fallback")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError:
                if path == LIB_PATH:
                    raise
                continue             # an older build under S1S2_LIB may lack the newest debug entries
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc, handle=None):
    if rc != OK:
        L = lib()
        msg = (L.s1s2_last_error(handle) if handle else L.s1s2_global_error()) or b""
        raise S1S2Error(f"libs1s2_b200 error {rc}: {msg.decode(errors='replace')}")
