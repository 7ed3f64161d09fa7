"""ctypes binding of libodpd.so (include/odpd.h).  Raw device pointers in, no torch types across the boundary.

Fails loudly when the library is missing: the product has no CPU path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libodpd.so")

CELLS = {"gru": 0, "lstm": 1, "dgru": 2, "deltagru": 3, "deltagru_tcnskip": 4, "pgjanet": 5, "dvrjanet": 6, "gmp": 7,
         "qgru": 8, "qgru_amp1": 9, "qgru_qat": 10, "qgru_amp1_qat": 11, "vdlstm": 12, "rvtdcnn": 13, "bojanet": 14, "tcnn": 15, "neuraltx": 16, "apnrru": 17, "mcldnn": 18, "deltajanet": 19, "deltagru_tcnskip_qat": 20}
F_NEED_DX, F_NEED_DW, F_SAVE, F_OVERWRITE_DW, F_ZERO_LOSS, F_X_BF16, F_TARGET_BF16 = 1, 2, 4, 8, 16, 32, 64


class OdpdDims(ctypes.Structure):
    _fields_ = [("cell", ctypes.c_int32), ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("H", ctypes.c_int32),
                ("K", ctypes.c_int32), ("flags", ctypes.c_uint32), ("thx", ctypes.c_float), ("thh", ctypes.c_float),
                ("tchunks", ctypes.c_int32), ("twarm", ctypes.c_int32), ("x_starts", ctypes.c_void_p), ("target_starts", ctypes.c_void_p)]


class OdpdError(RuntimeError):
    pass


_lib = None
_vp, _i64, _i32, _dbl, _flt = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double, ctypes.c_float
_DP = ctypes.POINTER(OdpdDims)

# every symbol include/odpd.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "odpd_version": (ctypes.c_int, []),
    "odpd_last_error": (ctypes.c_char_p, []),
    "odpd_n_params": (_i64, [_i32, _i32, _i32]),
    "odpd_saved_bytes": (_i64, [_DP]),
    "odpd_bwd_workspace_bytes": (_i64, [_DP]),
    "odpd_chunk_plan": (ctypes.c_int, [_DP, _i32, ctypes.POINTER(_i32)]),
    "odpd_chunk_plan_model": (ctypes.c_int, [_i32, _i32, _i32, _i32, _i32, _i32, ctypes.POINTER(_i32)]),
    "odpd_backbone_fwd": (ctypes.c_int, [_DP, _vp, _vp, _vp, _vp, _vp, _dbl, _vp, _vp, _vp]),
    "odpd_backbone_bwd": (ctypes.c_int, [_DP, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp, _vp]),
    "odpd_clip_adamw": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _flt, _flt, _flt, _flt, _flt, _vp, _vp,
                                       ctypes.c_int, _vp]),
    "odpd_fuse_next_bwd_with_adamw": (ctypes.c_int, [_vp, _vp, _vp, _vp, _flt, _flt, _flt, _flt, _flt, _vp, _vp, _vp]),
    "odpd_nmse_sums": (ctypes.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "odpd_dft_magnitude": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "odpd_dp_buffer_bytes": (_i64, [_i64]),
    "odpd_dp_alloc": (ctypes.c_int, [_i64, ctypes.POINTER(_vp)]),
    "odpd_dp_free": (ctypes.c_int, [_vp]),
    "odpd_dp_ipc_handle": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "odpd_dp_ipc_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "odpd_dp_ipc_close": (ctypes.c_int, [_vp]),
    "odpd_dp_publish_next_bwd": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, _i64, _vp, _vp]),
    "odpd_dp_clip_adamw": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _flt, _flt,
                                          _flt, _flt, _flt, _vp, _vp, _vp, _vp, _vp]),
}


def lib():
    """Load libodpd.so once. Raises OdpdError if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OdpdError(f"{LIB_PATH} is missing: build it with `make -C opendpd_b200/csrc` "
                            "(there is no CPU fallback for the native backbones)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise OdpdError(f"libodpd error {rc}: {lib().odpd_last_error().decode()}")


def n_params(cell, H, K=0):
    return int(lib().odpd_n_params(CELLS[cell], int(H), int(K)))
