"""sys.modules shims that let the UNMODIFIED reference scripts import and run on top of the
CUDA solver when the third-party packages they import are absent (as in this image):

    opengen              -> host.opengen_compat   (the solver boundary, src/path_generator.py:3)
    casadi.casadi        -> symbolic no-op stub   (only used by MpcModule.build(), src/mpc/mpc_generator.py:2)
    extremitypathfinder  -> host.planner          (src/visibility/visibility.py:1-3)
    pyclipper            -> host.planner.offset_polygon (src/visibility/visibility.py:5,43,90-105)
    matplotlib, cv2      -> inert stubs (plotting is out of scope)

`install()` never shadows a package that is really installed.
"""
import importlib
import importlib.util
import sys
import types
from unittest import mock

from . import opengen_compat, planner

CLIPPER_SCALE = 2 ** 31  # pyclipper.scale_to_clipper default


def _missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


# --- pyclipper ---------------------------------------------------------------------------
def _scale(value, f, rnd):
    if isinstance(value, (list, tuple)):
        return [_scale(v, f, rnd) for v in value]
    v = value * f
    return int(round(v)) if rnd else v


class _PyclipperOffset:
    def __init__(self, miter_limit=2.0, arc_tolerance=0.25):
        self.MiterLimit = miter_limit
        self._paths = []

    def Clear(self):
        self._paths = []

    def AddPath(self, path, join_type, end_type):
        self._paths.append([(p[0], p[1]) for p in path])

    def Execute(self, delta):
        out = []
        for path in self._paths:
            poly = planner.offset_polygon([(x / CLIPPER_SCALE, y / CLIPPER_SCALE) for x, y in path],
                                          delta / CLIPPER_SCALE, self.MiterLimit)
            out.append([[int(round(x * CLIPPER_SCALE)), int(round(y * CLIPPER_SCALE))] for x, y in poly])
        return out


def _make_pyclipper():
    m = types.ModuleType("pyclipper")
    m.JT_MITER, m.JT_ROUND, m.JT_SQUARE = 2, 1, 0
    m.ET_CLOSEDPOLYGON = 0
    m.PyclipperOffset = _PyclipperOffset
    m.scale_to_clipper = lambda v, scale=CLIPPER_SCALE: _scale(v, scale, True)
    m.scale_from_clipper = lambda v, scale=CLIPPER_SCALE: _scale(v, 1.0 / scale, False)
    return m


# --- extremitypathfinder --------------------------------------------------------------------
def _make_epf():
    pkg = types.ModuleType("extremitypathfinder")
    core = types.ModuleType("extremitypathfinder.extremitypathfinder")
    plot = types.ModuleType("extremitypathfinder.plotting")
    core.PolygonEnvironment = planner.PolygonEnvironment
    pkg.PolygonEnvironment = planner.PolygonEnvironment

    class PlottingEnvironment(planner.PolygonEnvironment):
        def __init__(self, plotting_dir=None):
            super().__init__()

    plot.PlottingEnvironment = PlottingEnvironment
    plot.draw_prepared_map = lambda *a, **k: None
    pkg.extremitypathfinder = core
    pkg.plotting = plot
    return {"extremitypathfinder": pkg, "extremitypathfinder.extremitypathfinder": core,
            "extremitypathfinder.plotting": plot}


# --- casadi: enough for `import casadi.casadi as cs` and for build() to run symbolically -------
def _make_casadi():
    pkg = types.ModuleType("casadi")
    core = mock.MagicMock(name="casadi.casadi")
    pkg.casadi = core
    return {"casadi": pkg, "casadi.casadi": core}


def _make_inert(names):
    out = {}
    for n in names:
        out[n] = mock.MagicMock(name=n)
    return out


def install(reference_config=None, solver_config=None, device=0, force=()):
    """Install the shims (only for packages that are missing, or listed in `force`) and bind
    the opengen shim to the given problem configuration.  Returns the names installed."""
    opengen_compat.configure(reference_config, solver_config, device)
    installed = []

    def put(mods, probe):
        if probe in force or _missing(probe):
            for k, v in mods.items():
                sys.modules[k] = v
            installed.append(probe)

    og = types.ModuleType("opengen")
    for k in ("tcp", "builder", "config", "constraints"):
        setattr(og, k, getattr(opengen_compat, k))
    put({"opengen": og, "opengen.tcp": og.tcp, "opengen.builder": og.builder, "opengen.config": og.config,
         "opengen.constraints": og.constraints}, "opengen")
    put(_make_casadi(), "casadi")
    put(_make_epf(), "extremitypathfinder")
    put({"pyclipper": _make_pyclipper()}, "pyclipper")
    put(_make_inert(["matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.lines",
                     "matplotlib.patches", "matplotlib.cm", "matplotlib.collections"]), "matplotlib")
    put(_make_inert(["cv2"]), "cv2")
    return installed
