"""ctypes loader for oracle/libodpd_oracle.so (test infrastructure only; never imported by opendpd_b200)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CELLS = {"gru": 0, "lstm": 1, "dgru": 2, "deltagru": 3, "deltagru_tcnskip": 4, "tres": 4, "pgjanet": 5,
         "dvrjanet": 6, "gmp": 7, "qgru": 8, "qgru_amp1": 9, "qgru_qat": 10, "qgru_amp1_qat": 11, "rvtdcnn": 13, "bojanet": 14, "tcnn": 15, "neuraltx": 16, "apnrru": 17, "mcldnn": 18, "deltajanet": 19, "deltagru_tcnskip_qat": 20, "tres_qat": 20}
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libodpd_oracle.so")
        if not os.path.exists(so):
            build()
        _lib = ctypes.CDLL(so)
        for f in (_lib.odpd_oracle_run_f32, _lib.odpd_oracle_run_f64):
            f.restype = ctypes.c_int
        for f in (_lib.odpd_oracle_set_dh_margin_f32, _lib.odpd_oracle_set_dh_margin_f64):
            f.restype, f.argtypes = None, [ctypes.c_void_p, ctypes.c_void_p]
        _lib.odpd_oracle_n_params.restype = ctypes.c_size_t
        _lib.odpd_oracle_n_params.argtypes = [ctypes.c_int] * 3
    return _lib


def n_params(cell, H, K=3):
    return int(lib().odpd_oracle_n_params(CELLS[cell], int(H), int(K)))


def run(cell, x, params, target=None, gout=None, H=0, K=3, thx=0.0, thh=0.0, want_grads=True, dtype=np.float32,
        loss_count=None, nthreads=1, want_masks=False, want_dh_margin=False):
    """Forward (+MSE) (+backward) of one (B,T,2) batch on the CPU oracle. Returns a dict of numpy arrays.
    want_dh_margin (delta cells, implies want_masks): also `dh_margin` (B,T,H) = |delta_h| - thh before masking."""
    want_masks = want_masks or want_dh_margin
    L = lib()
    dt = np.dtype(dtype)
    fn = L.odpd_oracle_run_f32 if dt == np.float32 else L.odpd_oracle_run_f64
    x = np.ascontiguousarray(x, dtype=dt)
    B, T = x.shape[0], x.shape[1]
    params = np.ascontiguousarray(params, dtype=dt)
    assert params.size == n_params(cell, H, K), (params.size, n_params(cell, H, K))
    out = np.zeros((B, T, 2), dtype=dt)
    tgt = None if target is None else np.ascontiguousarray(target, dtype=dt)
    go = None if gout is None else np.ascontiguousarray(gout, dtype=dt)
    bwd = want_grads and (tgt is not None or go is not None)
    gx = np.zeros((B, T, 2), dtype=dt) if bwd else None
    gp = np.zeros(params.size, dtype=dt) if bwd else None
    loss = ctypes.c_double(0.0)
    mx = np.zeros((B, T), dtype=np.uint64) if want_masks else None
    mh = np.zeros((B, T), dtype=np.uint64) if want_masks else None
    stats = np.zeros(4, dtype=np.int64)
    if loss_count is None:
        loss_count = float(2 * B * T)

    def p(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    margin = None
    setm = L.odpd_oracle_set_dh_margin_f32 if dt == np.float32 else L.odpd_oracle_set_dh_margin_f64
    if want_dh_margin:
        margin = np.zeros((B, T, int(H)), dtype=dt)
        setm(p(margin), p(mh))
    rc = fn(ctypes.c_int(CELLS[cell]), ctypes.c_int(B), ctypes.c_int(T), ctypes.c_int(int(H)), ctypes.c_int(int(K)),
            ctypes.c_double(float(thx)), ctypes.c_double(float(thh)), p(x), p(tgt), p(go), p(params), p(out),
            ctypes.byref(loss), ctypes.c_double(loss_count), p(gx), p(gp), p(mx), p(mh), p(stats),
            ctypes.c_int(int(nthreads)))
    if want_dh_margin:
        setm(None, None)
    if rc != 0:
        raise RuntimeError("oracle rejected the arguments")
    return dict(dh_margin=margin, out=out, loss=loss.value if tgt is not None else None, gx=gx, gparams=gp, mask_x=mx, mask_h=mh,
                stats=stats)
