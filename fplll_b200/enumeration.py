"""Host-side mirror of fplll's enumeration entry points over the C-ABI of include/b200enum.h.

`enumerate_svp` is the flattened form of the external-enumerator hook (fplll/enum/enumerate_ext_api.h:88-92);
`Enumeration` mirrors fplll/enum/enumerate.h:78-111 for the case BKZ uses (no target, no subtree, primal)."""
import ctypes as C

import numpy as np

from ._lib import B200Error, load

FIXED_RADIUS, DUAL, FINDSUBSOLS = 1, 2, 4
_P = C.POINTER
_CB = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double, _P(C.c_double))
_SUBCB = C.CFUNCTYPE(None, C.c_void_p, C.c_double, _P(C.c_double), C.c_int)
IPC_HANDLE_BYTES = 64


class _Stats(C.Structure):
    _fields_ = [("host_nodes", C.c_uint64), ("device_nodes", C.c_uint64), ("leaves", C.c_uint64),
                ("top_levels", C.c_int), ("n_roots", C.c_int), ("n_solutions", C.c_int), ("n_devices", C.c_int), ("n_rounds", C.c_int),
                ("final_maxdist", C.c_double), ("device_ms", C.c_float), ("host_breadth_us", C.c_float),
                ("total_us", C.c_float)]


_done = False


def _lib():
    global _done
    L = load("libb200enum.so")
    if not _done:
        L.b200enum_last_error.restype = C.c_char_p
        L.b200enum_run.argtypes = [C.c_int, C.c_double, _P(C.c_double), _P(C.c_double), _P(C.c_double), C.c_int,
                                   _P(C.c_int), C.c_int, C.c_int, C.c_int, _CB, C.c_void_p, _P(C.c_uint64),
                                   _P(_Stats)]
        L.b200enum_run_ex.argtypes = [C.c_int, C.c_double, _P(C.c_double), _P(C.c_double), _P(C.c_double), C.c_int,
                                      _P(C.c_int), C.c_int, C.c_int, C.c_int, _CB, _SUBCB, C.c_void_p, _P(C.c_uint64),
                                      _P(_Stats)]
        L.b200enum_ipc_export.argtypes = [C.c_int, C.c_char_p]
        L.b200enum_ipc_attach.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.b200enum_ipc_detach.argtypes = [C.c_int]
        _done = True
    return L


def ipc_export(device):
    """CUDA IPC handle (64 bytes) of this process's enumerator words on `device` (include/b200enum.h)."""
    buf = C.create_string_buffer(IPC_HANDLE_BYTES)
    rc = _lib().b200enum_ipc_export(int(device), buf)
    if rc != 0:
        raise B200Error("b200enum_ipc_export failed (%d): %s" % (rc, _lib().b200enum_last_error().decode()))
    return buf.raw


def ipc_attach(device, world, rank, handles):
    """handles: world * 64 bytes, rank-major (all-gathered ipc_export results)."""
    rc = _lib().b200enum_ipc_attach(int(device), int(world), int(rank), bytes(handles))
    if rc != 0:
        raise B200Error("b200enum_ipc_attach failed (%d): %s" % (rc, _lib().b200enum_last_error().decode()))


def ipc_detach(device):
    _lib().b200enum_ipc_detach(int(device))


def enumerate_svp(mut, rdiag, pruning, maxdist, fixed_radius=False, devices=None, shard=(0, 1), flags=0, dual=False,
                  findsubsols=False):
    """Runs the device enumerator.  mut[k, j] = mu(j, k) for j > k (the hook's transposed layout), everything
    normalised like ExternalEnumeration::enumerate does.  Returns dict(solutions=[(dist, x)...] in order of
    improvement, nodes[d], stats{...}, subsolutions={level: (dist, x)} when findsubsols); raises B200Error on failure
    (never falls back to a CPU enumerator).  dual=True: dual SVP of the block (enumerate.cpp:100-124), mut / rdiag are
    still the primal block's."""
    rdiag = np.ascontiguousarray(rdiag, np.float64)
    d = rdiag.shape[0]
    mut = np.ascontiguousarray(mut, np.float64).reshape(d, d)
    pr = None if pruning is None else np.ascontiguousarray(pruning, np.float64)
    sols = []

    def cb(ctx, dist, sol):
        sols.append((dist, np.array([sol[i] for i in range(d)])))
        return dist  # FastEvaluator(1): new bound = this solution (enum/evaluator.h:122-156)

    cbf = _CB(cb)
    subsols = {}

    def subcb(ctx, dist, sub, offset):
        subsols[int(offset)] = (dist, np.array([sub[i] for i in range(d)]))

    subf = _SUBCB(subcb)
    flags |= (DUAL if dual else 0) | (FINDSUBSOLS if findsubsols else 0)
    nodes = np.zeros(d, np.uint64)
    st = _Stats()
    dv = None
    nd = 0
    if devices is not None:
        dv = np.ascontiguousarray(devices, np.int32)
        nd = len(dv)
    rc = _lib().b200enum_run_ex(d, float(maxdist), mut.ctypes.data_as(_P(C.c_double)),
                             rdiag.ctypes.data_as(_P(C.c_double)),
                             pr.ctypes.data_as(_P(C.c_double)) if pr is not None else None,
                             flags | (FIXED_RADIUS if fixed_radius else 0),
                             dv.ctypes.data_as(_P(C.c_int)) if dv is not None else None, nd,
                             int(shard[0]), int(shard[1]), cbf, subf if findsubsols else _SUBCB(), None,
                             nodes.ctypes.data_as(_P(C.c_uint64)),
                             C.byref(st))
    if rc != 0:
        raise B200Error("b200enum_run failed (%d): %s" % (rc, _lib().b200enum_last_error().decode()))
    stats = {f: getattr(st, f) for f, _ in _Stats._fields_}
    return dict(solutions=sols, nodes=nodes, stats=stats, subsolutions=subsols)
