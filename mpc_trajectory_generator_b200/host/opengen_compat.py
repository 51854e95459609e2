"""`opengen`-shaped module over the CUDA solver: the drop-in for the reference's
`import opengen as og` on the solve path.

What the reference uses from opengen (and what answers here):
  og.tcp.OptimizerTcpManager(path)   src/path_generator.py:218-219  -> OptimizerTcpManager
     .start() .ping() .kill()        src/path_generator.py:220-222,408,417
     .call(parameters)               src/mpc/mpc_generator.py:206   -> SolverResponse
  response.is_ok() / .get()          src/mpc/mpc_generator.py:209-219
     ok:  .solution .exit_status .solve_time_ms      error: .code .message
  og.builder / og.config / og.constraints             src/mpc/mpc_generator.py:153,168,173-193
     -> OpEnOptimizerBuilder.build() compiles / loads libnmpc_b200.so instead of running
        CasADi codegen + cargo.

The sizes and the sampling time are build-time constants of the reference's generated
solver (configs/default.yaml:6); here they come from `configure(reference_config)`.
There is no CPU fallback: without the CUDA library / a device, start() raises.
"""
import types

import numpy as np

from ..solver import EXIT_STATUS_NAMES, NmpcConfig, NmpcSolver, load_library
from .. import _build

_ACTIVE = {"cfg": None, "device": 0}


def configure(reference_config=None, solver_config=None, device=0, max_duration_micros=0):
    """Bind the module to a problem size: either the reference's config dotdict
    (src/utils/config.py:52-72) or an explicit NmpcConfig.  `max_duration_micros` > 0 turns on the solver's time
    budget (the reference builds its solver with 500 000, src/mpc/mpc_generator.py:9,186); 0 = iteration caps only."""
    if solver_config is None:
        if reference_config is None:
            solver_config = NmpcConfig.default()
        else:
            solver_config = NmpcConfig.from_reference_config(reference_config)
        solver_config.max_duration_micros = int(max_duration_micros)
    _ACTIVE["cfg"] = solver_config
    _ACTIVE["device"] = int(device)
    return solver_config


class SolverStatus:
    """Fields of OpEn's TCP reply that a successful call exposes."""

    def __init__(self, u, status, stats, ms):
        self.exit_status = EXIT_STATUS_NAMES[int(status)]
        self.num_outer_iterations = int(stats["outer_iterations"])
        self.num_inner_iterations = int(stats["inner_iterations"])
        self.last_problem_norm_fpr = float(stats["last_norm_fpr"])
        self.f1_infeasibility = float(stats["delta_y_norm_over_c"])
        self.f2_norm = float(stats["f2_norm"])
        self.penalty = float(stats["penalty"])
        self.cost = float(stats["cost"])
        self.solve_time_ms = float(ms)          # device time of the solve kernel for this call
        self.solution = [float(x) for x in u]
        self.lagrange_multipliers = []


class SolverError:
    def __init__(self, code, message):
        self.code = code
        self.message = message


class SolverResponse:
    def __init__(self, payload, ok):
        self._payload, self._ok = payload, ok

    def is_ok(self):
        return self._ok

    def get(self):
        return self._payload


class OptimizerTcpManager:
    """Duck type of og.tcp.OptimizerTcpManager backed by one NmpcSolver (one CUDA device).
    State that OpEn's server keeps between requests (the decision vector and the
    multipliers, because the reference sends only `p`) lives in the solver handle."""

    def __init__(self, optimizer_path=None, ip=None, port=None, solver_config=None, device=None):
        self._path = optimizer_path
        self._cfg = solver_config or _ACTIVE["cfg"] or NmpcConfig.default()
        self._device = _ACTIVE["device"] if device is None else device
        self._solver = None

    def start(self):
        if self._solver is None:
            self._solver = NmpcSolver(self._cfg, self._device)

    def ping(self):
        if self._solver is None:
            raise RuntimeError("solver not started")
        self._solver.ping()
        return {"Pong": 1}

    def kill(self):
        if self._solver is not None:
            self._solver.close()
            self._solver = None

    def call(self, p, initial_guess=None, initial_y=None, initial_penalty=None, buffer_len=4096, max_data_size=1048576):
        if self._solver is None:
            raise RuntimeError("solver not started")
        if len(p) != self._solver.np:
            return SolverResponse(SolverError(3003, "wrong number of parameters"), False)
        if initial_guess is not None or initial_y is not None or initial_penalty is not None:
            return SolverResponse(SolverError(1600, "initial guess / multipliers / penalty are not used by the "
                                                    "reference (src/mpc/mpc_generator.py:206) and not supported"), False)
        u, status, stats, ms = self._solver.call(np.asarray(p, dtype=np.float64))
        if int(status) == 3:
            return SolverResponse(SolverError(2000, "Problem solution failed (solver error)"), False)
        return SolverResponse(SolverStatus(u, status, stats, ms), True)

    # batched extension (not in opengen): B independent problems in one launch
    def call_batch(self, P, U0=None, Y0=None):
        if self._solver is None:
            raise RuntimeError("solver not started")
        return self._solver.solve_batch(P, U0, Y0)


# --- og.builder / og.config / og.constraints as used by MpcModule.build() -----------------
class Rectangle:
    def __init__(self, xmin, xmax):
        self.xmin, self.xmax = xmin, xmax


class Problem:
    def __init__(self, u, p, cost):
        self.u, self.p, self.cost = u, p, cost
        self.penalty = self.bounds = self.alm = self.alm_set = None

    def with_penalty_constraints(self, f2):
        self.penalty = f2
        return self

    def with_constraints(self, bounds):
        self.bounds = bounds
        return self

    def with_aug_lagrangian_constraints(self, f1, set_c, set_y=None):
        self.alm, self.alm_set = f1, set_c
        return self


class _Chain:
    """config objects whose with_*() calls chain (BuildConfiguration, OptimizerMeta, SolverConfiguration)."""

    def __init__(self):
        self.settings = {}

    def __getattr__(self, name):
        if name.startswith("with_"):
            def setter(*args):
                self.settings[name[5:]] = args[0] if len(args) == 1 else args
                return self
            return setter
        raise AttributeError(name)


class OpEnOptimizerBuilder:
    """build() = compile (nvcc, sm_100a) or load the prebuilt CUDA library; the symbolic
    problem passed by MpcModule.build() is only sanity-checked against the configured sizes."""

    def __init__(self, problem, metadata=None, build_configuration=None, solver_configuration=None):
        self.problem, self.meta, self.build_cfg, self.solver_cfg = problem, metadata, build_configuration, solver_configuration

    def with_verbosity_level(self, level):
        return self

    def build(self):
        cfg = _ACTIVE["cfg"] or NmpcConfig.default()
        b = self.problem.bounds
        if b is not None and hasattr(b.xmin, "__len__") and len(b.xmin) != 2 * cfg.N_hor:
            raise ValueError(f"problem has {len(b.xmin)} decision variables, configured horizon expects {2 * cfg.N_hor}")
        tol = getattr(self.solver_cfg, "settings", {}).get("tolerance") if self.solver_cfg else None
        if tol is not None:
            cfg.tolerance = float(tol)
        if _build.find_nvcc():
            _build.build_library()
        load_library()
        return {"library": _build.LIB}


tcp = types.SimpleNamespace(OptimizerTcpManager=OptimizerTcpManager)
builder = types.SimpleNamespace(Problem=Problem, OpEnOptimizerBuilder=OpEnOptimizerBuilder)
config = types.SimpleNamespace(BuildConfiguration=_Chain, OptimizerMeta=_Chain, SolverConfiguration=_Chain)
constraints = types.SimpleNamespace(Rectangle=Rectangle)
