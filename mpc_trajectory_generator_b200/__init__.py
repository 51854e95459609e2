"""B200-native batched NMPC solver behind the solver-call surface of
wljungbergh/mpc-trajectory-generator (mng.call, src/mpc/mpc_generator.py:206)."""
from .solver import NmpcConfig, NmpcSolver, NmpcError, EXIT_STATUS_NAMES, STATS_DTYPE, param_len  # noqa: F401

from .fleet import FleetPlan, NmpcFleet  # noqa: F401,E402

__all__ = ["FleetPlan", "NmpcFleet", "NmpcConfig", "NmpcSolver", "NmpcError", "EXIT_STATUS_NAMES", "STATS_DTYPE", "param_len"]
