"""ctypes loader for oracle/libnmpc_oracle.so (the C restatement in nmpc_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnmpc_oracle.so")
_SO_SERIAL = os.path.join(_HERE, "libnmpc_oracle_serial.so")   # reference arithmetic (libm, divisions, serial sums)
_SO_NATIVE = os.path.join(_HERE, "_native", "libnmpc_oracle.so")  # -march=native copy for the CPU arm of bench.py


class Config(C.Structure):
    """struct nmpc_config (include/nmpc_b200.h)."""
    _fields_ = [
        ("N_hor", C.c_int32), ("Nobs", C.c_int32), ("Ndynobs", C.c_int32),
        ("lbfgs_memory", C.c_int32), ("max_inner_iterations", C.c_int32),
        ("max_outer_iterations", C.c_int32), ("max_duration_micros", C.c_int32), ("reserved1", C.c_int32),
        ("ts", C.c_double),
        ("lin_vel_min", C.c_double), ("lin_vel_max", C.c_double), ("ang_vel_max", C.c_double),
        ("lin_acc_min", C.c_double), ("lin_acc_max", C.c_double), ("ang_acc_max", C.c_double),
        ("tolerance", C.c_double), ("initial_tolerance", C.c_double), ("delta_tolerance", C.c_double),
        ("inner_tolerance_update", C.c_double), ("penalty_update_factor", C.c_double),
        ("initial_penalty", C.c_double), ("sufficient_decrease_coeff", C.c_double),
    ]


STATS_DTYPE = np.dtype([
    ("exit_status", np.int32), ("outer_iterations", np.int32), ("inner_iterations", np.int32),
    ("n_cost_evals", np.int32), ("n_grad_evals", np.int32), ("reserved", np.int32),
    ("last_norm_fpr", np.float64), ("delta_y_norm_over_c", np.float64), ("f2_norm", np.float64),
    ("penalty", np.float64), ("cost", np.float64),
])
assert STATS_DTYPE.itemsize == 64


def build(force=False):
    """Compile both builds of the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "nmpc_oracle.c")
    for so in (_SO, _SO_SERIAL):
        if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-B", os.path.basename(so)], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
    return _SO


def use_native():
    """bench.py's CPU arm: rebuild the contract oracle with -march=native ON THIS HOST (oracle/_native/) and
    load that copy from now on.  Same source, same results (no implicit contraction, no reassociation)."""
    global _lib
    try:
        subprocess.check_call(["make", "-C", _HERE, "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        return False
    _lib = _load(_SO_NATIVE)
    return True


_lib = None
_lib_serial = None


def _load(path):
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.nmpc_oracle_solve_batch.argtypes = [C.POINTER(Config), C.c_int32, dp, dp, dp,
                                              C.POINTER(C.c_int32), C.c_void_p, C.c_int]
        L.nmpc_oracle_solve_batch.restype = C.c_int
        L.nmpc_oracle_eval_batch.argtypes = [C.POINTER(Config), C.c_int32, dp, dp, dp, dp, dp, dp, dp, dp]
        L.nmpc_oracle_eval_batch.restype = C.c_int
        L.nmpc_oracle_sincos.argtypes = [C.c_double, dp, dp]
        L.nmpc_oracle_max_threads.restype = C.c_int
        L.nmpc_default_config.argtypes = [C.POINTER(Config)]
        L.nmpc_param_len.argtypes = [C.POINTER(Config)]
        L.nmpc_param_len.restype = C.c_int32
        L.nmpc_oracle_layout.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.nmpc_oracle_is_serial.restype = C.c_int
        return L


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = _load(_SO)
    return _lib


def lib_serial():
    """The -DNMPC_ORACLE_SERIAL build: reference arithmetic, independent of the kernel's contract."""
    global _lib_serial
    if _lib_serial is None:
        if not os.path.exists(_SO_SERIAL):
            build()
        _lib_serial = _load(_SO_SERIAL)
        assert _lib_serial.nmpc_oracle_is_serial() == 1
    return _lib_serial


def layout(N):
    """(G, S): lanes per evaluation group and consecutive horizon steps per lane (the kernel's layout rule)."""
    g, s = C.c_int(), C.c_int()
    lib().nmpc_oracle_layout(int(N), C.byref(g), C.byref(s))
    return g.value, s.value


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def default_config(**kw):
    cfg = Config()
    lib().nmpc_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def param_len(cfg):
    return int(lib().nmpc_param_len(C.byref(cfg)))


def solve_batch(cfg, P, U0=None, Y0=None, nthreads=0, serial=False):
    """-> (U, Y, status, stats).  U0/Y0 None = zeros (a freshly started server).
    serial=True: the reference-arithmetic build (libm, divisions, serial sums)."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    B = P.shape[0]
    n2 = 2 * cfg.N_hor
    assert P.shape[1] == param_len(cfg)
    U = np.zeros((B, n2)) if U0 is None else np.array(U0, dtype=np.float64, order="C").reshape(B, n2)
    Y = np.zeros((B, n2)) if Y0 is None else np.array(Y0, dtype=np.float64, order="C").reshape(B, n2)
    status = np.zeros(B, dtype=np.int32)
    stats = np.zeros(B, dtype=STATS_DTYPE)
    rc = (lib_serial() if serial else lib()).nmpc_oracle_solve_batch(C.byref(cfg), B, _dp(P), _dp(U), _dp(Y),
                                       status.ctypes.data_as(C.POINTER(C.c_int32)),
                                       stats.ctypes.data_as(C.c_void_p), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle solve failed rc={rc}")
    return U, Y, status, stats


def eval_batch(cfg, P, U, c, Y=None, serial=False):
    """-> (psi[B], grad[B,2N], F1[B,2N], F2[B,Nobs+Ndynobs])."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    B = P.shape[0]
    n2 = 2 * cfg.N_hor
    c = np.ascontiguousarray(np.broadcast_to(np.asarray(c, dtype=np.float64), (B,)))
    Yc = None if Y is None else np.ascontiguousarray(Y, dtype=np.float64)
    psi = np.zeros(B)
    grad = np.zeros((B, n2))
    F1 = np.zeros((B, n2))
    F2 = np.zeros((B, cfg.Nobs + cfg.Ndynobs))
    rc = (lib_serial() if serial else lib()).nmpc_oracle_eval_batch(C.byref(cfg), B, _dp(P), _dp(U), _dp(c), _dp(Yc),
                                      _dp(psi), _dp(grad), _dp(F1), _dp(F2))
    if rc != 0:
        raise RuntimeError(f"oracle eval failed rc={rc}")
    return psi, grad, F1, F2


def sincos(x):
    s = C.c_double()
    c = C.c_double()
    lib().nmpc_oracle_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def max_threads():
    return int(lib().nmpc_oracle_max_threads())
