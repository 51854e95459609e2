"""ctypes binding of the C ABI in include/nmpc_b200.h (libnmpc_b200.so).

Host-side mirror of the reference's solver-call surface: `NmpcSolver.call(p)` is what
`mng.call(parameters)` does for one problem (src/mpc/mpc_generator.py:206), with the
server-side warm start the reference relies on; `solve_batch` is the same solve over a
batch of independent problems.  There is no CPU fallback: if the CUDA library cannot be
loaded or no device answers, construction raises NmpcError.
"""
import ctypes as C
import os

import numpy as np

from . import _build

EXIT_STATUS_NAMES = {0: "Converged", 1: "NotConvergedIterations", 2: "NotConvergedOutOfTime",
                     3: "NotFiniteComputation"}
NZ = 20


class NmpcError(RuntimeError):
    pass


class NmpcConfig(C.Structure):
    """struct nmpc_config.  Sizes and bounds are the reference's build-time constants
    (configs/default.yaml:6-13,18,38-39); the rest are opengen 0.6.4 solver defaults
    (src/mpc/mpc_generator.py:184-186 sets only the tolerance)."""
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

    @classmethod
    def default(cls, **overrides):
        """configs/default.yaml + opengen defaults (same values as nmpc_default_config)."""
        cfg = cls(N_hor=20, Nobs=10, Ndynobs=3, lbfgs_memory=10, max_inner_iterations=500,
                  max_outer_iterations=10, ts=0.2, lin_vel_min=-0.5, lin_vel_max=1.5, ang_vel_max=0.5,
                  lin_acc_min=-1.0, lin_acc_max=1.0, ang_acc_max=3.0, tolerance=1e-4,
                  initial_tolerance=1e-4, delta_tolerance=1e-4, inner_tolerance_update=0.1,
                  penalty_update_factor=5.0, initial_penalty=1.0, sufficient_decrease_coeff=0.1)
        for k, v in overrides.items():
            if k not in dict(cls._fields_):
                raise KeyError(k)
            setattr(cfg, k, v)
        return cfg

    @classmethod
    def from_reference_config(cls, config, **overrides):
        """From the reference's `Configurator(...).configurate()` dotdict (src/utils/config.py:52-72)."""
        return cls.default(N_hor=int(config.N_hor), Nobs=int(config.Nobs), Ndynobs=int(config.Ndynobs),
                           ts=float(config.ts), lin_vel_min=float(config.lin_vel_min),
                           lin_vel_max=float(config.lin_vel_max), ang_vel_max=float(config.ang_vel_max),
                           lin_acc_min=float(config.lin_acc_min), lin_acc_max=float(config.lin_acc_max),
                           ang_acc_max=float(config.ang_acc_max), **overrides)

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


STATS_DTYPE = np.dtype([
    ("exit_status", np.int32), ("outer_iterations", np.int32), ("inner_iterations", np.int32),
    ("n_cost_evals", np.int32), ("n_grad_evals", np.int32), ("reserved", np.int32),
    ("last_norm_fpr", np.float64), ("delta_y_norm_over_c", np.float64), ("f2_norm", np.float64),
    ("penalty", np.float64), ("cost", np.float64),
])
assert STATS_DTYPE.itemsize == 64


def param_len(cfg):
    """nz + N + 3*Nobs + 5*Ndynobs*N + 3*N  (src/mpc/mpc_generator.py:71)."""
    return NZ + cfg.N_hor + 3 * cfg.Nobs + 5 * cfg.Ndynobs * cfg.N_hor + 3 * cfg.N_hor


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def load_library(build=True):
    """dlopen libnmpc_b200.so (building it with nvcc first if it is missing/stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("NMPC_B200_LIB")   # tools/: a tuning/profiling build of the same sources
    if path is None:
        path = _build.LIB
        if build and _build.is_stale() and _build.find_nvcc():
            _build.build_library()
    if not os.path.exists(path):
        raise NmpcError(f"{path} not found and nvcc unavailable: the CUDA solver library is required "
                        "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.nmpc_default_config.argtypes = [C.POINTER(NmpcConfig)]
    L.nmpc_param_len.argtypes = [C.POINTER(NmpcConfig)]
    L.nmpc_param_len.restype = C.c_int32
    L.nmpc_create.argtypes = [C.POINTER(NmpcConfig), C.c_int, C.POINTER(vp)]
    L.nmpc_destroy.argtypes = [vp]
    L.nmpc_ping.argtypes = [vp]
    L.nmpc_call.argtypes = [vp, _dp, _dp, _ip, vp]
    L.nmpc_reset_warm_start.argtypes = [vp]
    L.nmpc_solve_batch.argtypes = [vp, C.c_int32, _dp, _dp, _dp, _ip, vp]
    L.nmpc_solve_batch_device.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.nmpc_eval_batch.argtypes = [vp, C.c_int32, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    L.nmpc_launch_count.argtypes = [vp]
    L.nmpc_launch_count.restype = C.c_int64
    L.nmpc_last_kernel_ms.argtypes = [vp]
    L.nmpc_last_kernel_ms.restype = C.c_double
    L.nmpc_last_error.argtypes = [vp]
    L.nmpc_last_error.restype = C.c_char_p
    L.nmpc_exit_status_name.argtypes = [C.c_int32]
    L.nmpc_exit_status_name.restype = C.c_char_p
    L.nmpc_abi_version.restype = C.c_int32
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


class NmpcSolver:
    """One solver instance bound to one CUDA device (≈ one running OpEn server)."""

    def __init__(self, cfg=None, device=0):
        self.cfg = cfg if cfg is not None else NmpcConfig.default()
        self.device = int(device)
        self.np = param_len(self.cfg)
        self.n2 = 2 * self.cfg.N_hor
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.nmpc_create(C.byref(self.cfg), self.device, C.byref(h))
        if rc != 0 or not h:
            raise NmpcError(f"nmpc_create failed (rc={rc}): no usable CUDA device {device} or unsupported "
                            f"config; this solver has no CPU fallback")
        self._h = h
        self._fleets = []   # weak references to NmpcFleet objects created on this handle (closed before it)

    # -- lifecycle (mng.start/ping/kill, src/path_generator.py:220-222,408,417) ----------
    def ping(self):
        self._check(self._lib.nmpc_ping(self._h), "ping")
        return True

    def close(self):
        if getattr(self, "_h", None):
            for ref in getattr(self, "_fleets", []):
                f = ref()
                if f is not None:
                    f.close()
            self._lib.nmpc_destroy(self._h)
            self._h = None

    kill = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.nmpc_last_error(self._h)
            raise NmpcError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")

    # -- single problem with server-side warm start (mng.call) ---------------------------
    def call(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64).ravel()
        if p.size != self.np:
            raise NmpcError(f"wrong number of parameters: got {p.size}, expected {self.np}")
        u = np.zeros(self.n2)
        st = C.c_int32()
        stats = np.zeros(1, dtype=STATS_DTYPE)
        self._check(self._lib.nmpc_call(self._h, _ptr(p), _ptr(u), C.byref(st), stats.ctypes.data_as(C.c_void_p)),
                    "nmpc_call")
        return u, int(st.value), stats[0], float(self._lib.nmpc_last_kernel_ms(self._h))

    def reset_warm_start(self):
        self._check(self._lib.nmpc_reset_warm_start(self._h), "reset_warm_start")

    # -- batch, host buffers ---------------------------------------------------------------
    def solve_batch(self, P, U0=None, Y0=None, want_stats=True):
        """-> (U[B,2N], Y[B,2N], status[B] int32, stats[B]).  U0/Y0 None = zeros (cold start)."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        if P.ndim != 2 or P.shape[1] != self.np:
            raise NmpcError(f"P must be [B, {self.np}]")
        B = P.shape[0]
        U = np.zeros((B, self.n2)) if U0 is None else np.array(U0, dtype=np.float64, order="C").reshape(B, self.n2)
        Y = np.zeros((B, self.n2)) if Y0 is None else np.array(Y0, dtype=np.float64, order="C").reshape(B, self.n2)
        status = np.zeros(B, dtype=np.int32)
        stats = np.zeros(B, dtype=STATS_DTYPE) if want_stats else None
        self._check(self._lib.nmpc_solve_batch(self._h, B, _ptr(P), _ptr(U), _ptr(Y), status.ctypes.data_as(_ip),
                                               None if stats is None else stats.ctypes.data_as(C.c_void_p)),
                    "nmpc_solve_batch")
        return U, Y, status, stats

    def solve_batch_into(self, P, U, Y, status, stats=None):
        """In-place variant for caller-owned (e.g. pinned) host buffers: U/Y are in/out
        (initial guess / multipliers in, solution / multiplier state out); no copies here."""
        B = P.shape[0]
        for a, shape, dt in ((P, (B, self.np), np.float64), (U, (B, self.n2), np.float64),
                             (Y, (B, self.n2), np.float64), (status, (B,), np.int32)):
            if a.shape != shape or a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise NmpcError(f"buffer must be C-contiguous {dt.__name__}{list(shape)}")
        if stats is not None and (stats.dtype != STATS_DTYPE or stats.shape != (B,)):
            raise NmpcError("stats must be STATS_DTYPE[B]")
        self._check(self._lib.nmpc_solve_batch(self._h, B, _ptr(P), _ptr(U), _ptr(Y), status.ctypes.data_as(_ip),
                                               None if stats is None else stats.ctypes.data_as(C.c_void_p)),
                    "nmpc_solve_batch")

    # -- batch, device buffers (raw pointers, e.g. torch tensors' data_ptr()) ---------------
    def solve_batch_device(self, B, dP, dU, dY=0, dstatus=0, dstats=0, stream=0):
        self._check(self._lib.nmpc_solve_batch_device(self._h, int(B), dP, dU, dY or None, dstatus or None,
                                                      dstats or None, stream or None), "nmpc_solve_batch_device")

    # -- parity hook -------------------------------------------------------------------------
    def eval_batch(self, P, U, c, Y=None):
        """-> (psi[B], grad[B,2N], F1[B,2N], F2[B,Nobs+Ndynobs])."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        U = np.ascontiguousarray(U, dtype=np.float64)
        B = P.shape[0]
        c = np.ascontiguousarray(np.broadcast_to(np.asarray(c, dtype=np.float64), (B,)))
        Yc = None if Y is None else np.ascontiguousarray(Y, dtype=np.float64)
        psi = np.zeros(B)
        grad = np.zeros((B, self.n2))
        F1 = np.zeros((B, self.n2))
        F2 = np.zeros((B, self.cfg.Nobs + self.cfg.Ndynobs))
        self._check(self._lib.nmpc_eval_batch(self._h, B, _ptr(P), _ptr(U), _ptr(c), _ptr(Yc), _ptr(psi),
                                              _ptr(grad), _ptr(F1), _ptr(F2)), "nmpc_eval_batch")
        return psi, grad, F1, F2

    @property
    def launch_count(self):
        return int(self._lib.nmpc_launch_count(self._h))

    @property
    def last_kernel_ms(self):
        return float(self._lib.nmpc_last_kernel_ms(self._h))
