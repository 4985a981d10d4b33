"""ctypes binding of libcamkifu_b200.so (the C ABI declared in include/camkifu_b200.h).

There is no CPU fallback: if the CUDA library is missing the import of this module's `lib()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CAMKIFU_B200_LIB: an instrumented build of the same library (tools/km_probe.py); never a different implementation
LIB_PATH = os.environ.get("CAMKIFU_B200_LIB") or os.path.join(HERE, "libcamkifu_b200.so")

CKB_OK = 0
_lib = None

# name -> (restype, argtypes); mirrors include/camkifu_b200.h one to one
_u8p, _f32p, _f64p, _i32p, _u64p, _vp = (C.c_void_p,) * 6
SIGNATURES = {
    "ckb_version": (C.c_int, []),
    "ckb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "ckb_destroy": (C.c_int, [C.c_void_p]),
    "ckb_last_error": (C.c_char_p, [C.c_void_p]),
    "ckb_rng_seed": (C.c_uint64, [C.c_uint32]),
    "ckb_rng_advance": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "ckb_invert_homography": (C.c_int, [_f64p, _f64p]),
    "ckb_zone_rects": (C.c_int, [C.c_int, _i32p]),
    "ckb_zone_mask": (C.c_int, [C.c_int, _u8p]),
    "ckb_warp": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _f64p, C.c_int, _u8p,
                           C.c_void_p]),
    "ckb_accumulate": (C.c_int, [C.c_void_p, _u8p, C.c_int, _f32p, C.c_float, C.c_int, _f32p, C.c_int, C.c_int,
                                 C.c_void_p]),
    "ckb_mog2_state_bytes": (C.c_size_t, [C.c_void_p]),
    "ckb_mog2_reset": (C.c_int, [C.c_void_p, _vp, C.c_void_p]),
    "ckb_mog2_apply": (C.c_int, [C.c_void_p, _u8p, C.c_int, _vp, C.c_longlong, C.c_void_p, _u8p, C.c_void_p]),
    "ckb_zone_fg_counts": (C.c_int, [C.c_void_p, _u8p, C.c_int, _vp, C.c_void_p]),
    "ckb_find_stones_workspace": (C.c_size_t, [C.c_void_p, C.c_int]),
    "ckb_find_stones": (C.c_int, [C.c_void_p, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u64p, _vp,
                                  C.c_size_t, _u8p, _u8p, _u8p, _f32p, _f64p, _i32p, C.c_void_p]),
    "ckb_find_stones_regions_workspace": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "ckb_find_stones_regions": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, _i32p, _u64p, _vp, C.c_size_t, _u8p, _u8p, _u8p,
                                          _f32p, _f64p, C.c_void_p]),
    "ckb_zone_means": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_void_p]),
    "ckb_history_vote": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, _u8p, C.c_void_p]),
    "ckb_set_cnn_weights": (C.c_int, [C.c_void_p, _f32p, C.c_size_t]),
    "ckb_cnn_workspace": (C.c_size_t, [C.c_void_p, C.c_int]),
    "ckb_cnn_forward": (C.c_int, [C.c_void_p, _u8p, C.c_int, _vp, C.c_size_t, _f32p, _u8p, _f32p, _u8p, C.c_void_p]),
    "ckb_cnn_forward_simt": (C.c_int, [C.c_void_p, _u8p, C.c_int, _vp, C.c_size_t, _f32p, _u8p, _f32p, _u8p,
                                       C.c_void_p]),
    "ckb_cnn_workspace_simt": (C.c_size_t, [C.c_void_p, C.c_int]),
    "ckb_cnn_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "ckb_cnn_debug_activation": (C.c_int, [C.c_void_p, _vp, C.c_int, C.c_int, _f32p, C.c_void_p]),
    "ckb_frame_roi": (C.c_int, [_f64p, C.c_int, C.c_int, C.c_int, _i32p]),
    "ckb_upload_frames": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _i32p, _u8p,
                                    C.c_size_t, C.c_size_t, C.c_void_p]),
    "ckb_jpeg_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _u8p, C.c_size_t, C.c_size_t,
                                  C.c_int, C.c_int, C.c_void_p]),
    "ckb_jpeg_backend": (C.c_char_p, [C.c_void_p]),
    "ckb_profile_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "ckb_profile_end": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, _f32p, _i32p]),
    "ckb_launch_count": (C.c_uint64, [C.c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("camkifu_b200: %s is missing — build it with `python -m camkifu_b200.build` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class CkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("camkifu_b200 error %d: %s" % (code, msg))
        self.code = code
