"""The C-ABI library loads and exports every symbol include/nmpc_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import mpc_trajectory_generator_b200 as pkg
from mpc_trajectory_generator_b200 import _build, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nmpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nmpc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    lib_path = _build.build_library()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nmpc_b200.h but not exported"


def test_struct_layout_and_helpers():
    L = solver.load_library()
    assert L.nmpc_abi_version() == 1
    cfg = pkg.NmpcConfig()
    L.nmpc_default_config(ctypes.byref(cfg))
    ref = pkg.NmpcConfig.default()
    assert cfg.as_dict() == ref.as_dict()                         # python defaults == C defaults
    assert ctypes.sizeof(pkg.NmpcConfig) == 8 * 4 + 14 * 8
    assert L.nmpc_param_len(ctypes.byref(cfg)) == 430 == pkg.param_len(cfg)   # R^430 at N=20
    assert pkg.param_len(pkg.NmpcConfig.default(N_hor=40)) == 810
    assert L.nmpc_exit_status_name(1) == b"NotConvergedIterations"   # configs/default.yaml:49 bad_exit_codes
    assert L.nmpc_exit_status_name(0) == b"Converged"
    assert pkg.STATS_DTYPE.itemsize == 64


def test_no_cpu_fallback_without_device():
    """On a box without a CUDA device the product must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        pkg.NmpcSolver()
    except pkg.NmpcError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("NmpcSolver() must raise without a CUDA device")


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mpc_trajectory_generator_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libnmpc_oracle" not in txt, f
