"""Host-side mirror of MatHouseholder<Z_NR<long>, FP_NR<double>> (fplll/householder.h:38) over include/b200hh.h."""
import ctypes as C

import numpy as np

from ._lib import B200Error, load

HOUSEHOLDER_DEFAULT, HOUSEHOLDER_ROW_EXPO, HOUSEHOLDER_OP_FORCE_LONG = 0, 1, 4  # householder.h:26-32
_P = C.POINTER
_done = False


def _lib():
    global _done
    L = load("libb200hh.so")
    if not _done:
        vp, i = C.c_void_p, C.c_int
        L.b200hh_last_error.restype = C.c_char_p
        L.b200hh_create.argtypes = [_P(vp), i, i, i, i, i, i]
        L.b200hh_destroy.argtypes = [vp]
        L.b200hh_destroy.restype = None
        L.b200hh_set_basis.argtypes = [vp, _P(C.c_int64)]
        L.b200hh_get_basis.argtypes = [vp, _P(C.c_int64)]
        for f in ("refresh_R_bf", "refresh_R", "update_R_last", "recover_R"):
            getattr(L, "b200hh_" + f).argtypes = [vp, i]
        L.b200hh_update_R.argtypes = [vp, i, i]
        L.b200hh_swap.argtypes = [vp, i, i]
        L.b200hh_size_reduce.argtypes = [vp, i, i, i, _P(C.c_int)]
        L.b200hh_set_updated_R_false.argtypes = [vp]
        L.b200hh_get_state.argtypes = [vp] + [_P(C.c_double)] * 5 + [_P(C.c_int64)] * 2 + [_P(C.c_int)]
        L.b200hh_time_update_R.argtypes = [vp, i, i, _P(C.c_float)]
        L.b200hh_sync.argtypes = [vp]
        L.b200hh_hlll.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, _P(C.c_int), _P(C.c_uint64)]
        _done = True
    return L


def _ck(rc, what):
    if rc != 0:
        raise B200Error("%s failed (%d): %s" % (what, rc, _lib().b200hh_last_error().decode()))


class MatHouseholder:
    """MatHouseholder(b, flags): b (d, n) or (batch, d, n) int64; reference method names (householder.h)."""

    def __init__(self, b, flags=HOUSEHOLDER_ROW_EXPO | HOUSEHOLDER_OP_FORCE_LONG, device=0, keep_history=True):
        b = np.ascontiguousarray(b, dtype=np.int64)
        if b.ndim == 2:
            b = b[None]
        self.batch, self.d, self.n = b.shape
        self._h = C.c_void_p()
        _ck(_lib().b200hh_create(C.byref(self._h), self.batch, self.d, self.n, flags, device, 1 if keep_history else 0),
            "b200hh_create")
        _ck(_lib().b200hh_set_basis(self._h, b.ctypes.data_as(_P(C.c_int64))), "set_basis")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib().b200hh_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def refresh_R_bf(self, i):
        _ck(_lib().b200hh_refresh_R_bf(self._h, i), "refresh_R_bf")

    def refresh_R(self, i):
        _ck(_lib().b200hh_refresh_R(self._h, i), "refresh_R")

    def update_R(self, i, last_j=True):
        _ck(_lib().b200hh_update_R(self._h, i, 1 if last_j else 0), "update_R")

    def update_R_last(self, i):
        _ck(_lib().b200hh_update_R_last(self._h, i), "update_R_last")

    def size_reduce(self, k, end, start=0):
        red = np.zeros(self.batch, np.int32)
        _ck(_lib().b200hh_size_reduce(self._h, k, end, start, red.ctypes.data_as(_P(C.c_int))), "size_reduce")
        return red.astype(bool)

    def swap(self, i, j):
        _ck(_lib().b200hh_swap(self._h, i, j), "swap")

    def recover_R(self, i):
        _ck(_lib().b200hh_recover_R(self._h, i), "recover_R")

    def set_updated_R_false(self):
        _ck(_lib().b200hh_set_updated_R_false(self._h), "set_updated_R_false")

    def state(self):
        B, d, n = self.batch, self.d, self.n
        R, V, bf = (np.empty((B, d, n)) for _ in range(3))
        sg, nsb = np.empty((B, d)), np.empty((B, d))
        re, en = np.empty((B, d), np.int64), np.empty((B, d), np.int64)
        meta = np.empty((B, 3), np.int32)
        b = np.empty((B, d, n), np.int64)
        dp = lambda a: a.ctypes.data_as(_P(C.c_double))
        _ck(_lib().b200hh_get_state(self._h, dp(R), dp(V), dp(bf), dp(sg), dp(nsb), re.ctypes.data_as(_P(C.c_int64)),
                                    en.ctypes.data_as(_P(C.c_int64)), meta.ctypes.data_as(_P(C.c_int))), "get_state")
        _ck(_lib().b200hh_get_basis(self._h, b.ctypes.data_as(_P(C.c_int64))), "get_basis")
        return dict(R=R, V=V, bf=bf, sigma=sg, norm_square_b=nsb, row_expo=re, expo_norm_square_b=en, b=b,
                    n_known_rows=meta[:, 0].copy(), n_known_cols=meta[:, 1].copy(), updated_R=meta[:, 2].copy())

    def hlll(self, delta=0.99, eta=0.51, theta=0.001, c=0.1):
        """HLLLReduction(m, delta, eta, theta, c, flags).hlll() (hlll.cpp:25-171), whole loop on the device.
        Returns (status per lattice, main-loop iterations per lattice)."""
        st = np.zeros(self.batch, np.int32)
        it = np.zeros(self.batch, np.uint64)
        _ck(_lib().b200hh_hlll(self._h, delta, eta, theta, c, st.ctypes.data_as(_P(C.c_int)),
                               it.ctypes.data_as(_P(C.c_uint64))), "hlll")
        return st, it

    def get_basis(self):
        b = np.empty((self.batch, self.d, self.n), np.int64)
        _ck(_lib().b200hh_get_basis(self._h, b.ctypes.data_as(_P(C.c_int64))), "get_basis")
        return b

    def time_update_R(self, i, reps):
        ms = C.c_float()
        _ck(_lib().b200hh_time_update_R(self._h, i, reps, C.byref(ms)), "time_update_R")
        return ms.value


RED_HLLL_FAILURE, RED_HLLL_NORM_FAILURE, RED_HLLL_SR_FAILURE = 9, 10, 11  # defs.h:164-166


def hlll_reduction(b, delta=0.99, eta=0.51, theta=0.001, c=0.1, device=0):
    """hlll_reduction(ZZ_mat<long>& b, delta, eta, theta, c, HM_FAST, FT_DOUBLE) — fplll/wrapper.cpp:789-806 with
    defaults LLL_DEF_DELTA / LLL_DEF_ETA / HLLL_DEF_THETA / HLLL_DEF_C (defs.h:143-151).  b: (d, n) or (batch, d, n)
    int64.  Returns (reduced basis, status) — per lattice for a batch; the input array is not modified."""
    a = np.ascontiguousarray(b, dtype=np.int64)
    single = a.ndim == 2
    m = MatHouseholder(a, HOUSEHOLDER_ROW_EXPO | HOUSEHOLDER_OP_FORCE_LONG, device=device, keep_history=True)
    try:
        st, _ = m.hlll(delta, eta, theta, c)
        out = m.get_basis()
    finally:
        m.close()
    return (out[0], int(st[0])) if single else (out, st)
