"""TEST INFRASTRUCTURE ONLY - loads the *unmodified* reference hot path from /root/reference.

This runs only in the build container (``/root/reference`` does not exist on the GPU box).  It is used
by ``tests/golden/make_golden.py`` to generate the committed golden fixtures and by the optional
``-m "not gpu"`` test that pins ``oracle/femder_oracle.py`` against the live reference.

``import femder`` cannot work here: ``femder/__init__.py:7-23`` eagerly imports gmsh, meshio, matplotlib,
seaborn, plotly, tmm, pytta and pyMKL, none of which are installed.  So:

1. permissive stub modules are registered for the plotting / meshing imports;
2. a fake ``pyMKL`` is registered whose ``pardisoSolver(A, mtype)`` wraps
   ``scipy.sparse.linalg.splu`` (phase 12 -> factor, phase 33 -> solve, clear -> no-op).  pyMKL is a
   third-party, un-vendored, unpinned dependency (``README.md:28``, ``requirements.txt:4``); it is a direct
   solver, so SuperLU on the same ``G`` is the stand-in the north star itself names;
3. ``np.cfloat`` (removed in NumPy 2, used at ``FEM_3D.py:1266``) is aliased to ``np.complex128``;
4. ``femder`` is created as an empty package and ``utils``, ``fem_numerical``, ``controlsair``,
   ``BoundaryConditions`` and ``FEM_3D`` are loaded from their files, unmodified.

Nothing from the reference is copied; the files are executed where they lie.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("FEMDER_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "femder", "FEM_3D.py"))


class _Anything:
    """Attribute/call sink for plotting and meshing modules the hot path never touches."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []  # behaves as a package

    def _getattr(attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Anything()

    m.__getattr__ = _getattr
    sys.modules[name] = m
    return m


class _PardisoStandIn:
    """``pyMKL.pardisoSolver`` call surface used at ``FEM_3D.py:1315-1318`` over SuperLU."""

    def __init__(self, A, mtype=13):
        from scipy.sparse import csc_matrix
        self._A = csc_matrix(A).astype(np.complex128)
        self._lu = None

    def run_pardiso(self, phase, rhs=None):
        from scipy.sparse.linalg import splu
        if phase in (11, 12, 22):
            self._lu = splu(self._A)
            return None
        if phase == 33:
            return self._lu.solve(np.asarray(rhs, dtype=np.complex128).ravel())
        raise ValueError(f"phase {phase} not supported by the stand-in")

    def clear(self):
        self._lu = None


_loaded = None


def load_reference(disable_jit=False):
    """Return a namespace with the reference's ``FEM_3D``, ``fem_numerical``, ``controlsair`` and
    ``BoundaryConditions`` modules.  ``disable_jit`` must be decided before the first call (numba reads
    ``NUMBA_DISABLE_JIT`` at import); Tet10 and off-node ``evaluate`` need it."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if disable_jit:
        os.environ["NUMBA_DISABLE_JIT"] = "1"
    if not hasattr(np, "cfloat"):
        np.cfloat = np.complex128
    np.sctypeDict.setdefault("cfloat", np.complex128)   # int_tetra_4gauss asks for dtype='cfloat' (FEM_3D.py:775-776): a NumPy 1 alias
    import numpy.matlib  # noqa: F401  (FEM_3D.py uses np.matlib.repmat without importing it)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.colors", "seaborn",
                 "plotly", "plotly.io", "plotly.graph_objs", "plotly.graph_objects", "plotly.figure_factory",
                 "plotly.express", "plotly.subplots", "gmsh", "meshio", "tmm", "tmm.tmm", "pytta", "pymoo"):
        if name not in sys.modules:
            _stub(name)
    mkl = types.ModuleType("pyMKL")
    mkl.pardisoSolver = _PardisoStandIn
    sys.modules["pyMKL"] = mkl

    pkg = types.ModuleType("femder")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "femder")]
    sys.modules["femder"] = pkg

    def load(mod):
        path = os.path.join(REFERENCE_ROOT, "femder", mod + ".py")
        spec = importlib.util.spec_from_file_location("femder." + mod, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules["femder." + mod] = m
        spec.loader.exec_module(m)
        setattr(pkg, mod, m)
        return m

    ns = types.SimpleNamespace()
    ns.utils = load("utils")
    ns.fem_numerical = load("fem_numerical")
    ns.controlsair = load("controlsair")
    ns.BoundaryConditions = load("BoundaryConditions")
    pkg.p2SPL = None
    ns.FEM_3D = load("FEM_3D")
    pkg.p2SPL = ns.FEM_3D.p2SPL
    _loaded = ns
    return ns


class PlainSource:
    def __init__(self, coord, q):
        self.coord = np.atleast_2d(np.asarray(coord, dtype=np.float64))
        self.q = np.asarray(q, dtype=np.float64).reshape(-1, 1)
        self.wavetype = "spherical"


class PlainReceiver:
    def __init__(self, coord):
        self.coord = np.atleast_2d(np.asarray(coord, dtype=np.float64))


def run_reference_compute(grid, freq, mu_by_group, src_coord, src_q, rcv_coord, c0=343.0, rho0=1.21, v_by_group=None, fluid=None):
    """Run the reference's own ``FEM3D.compute()`` + ``evaluate()`` on ``grid``.  Returns the FEM3D object.
    ``fluid``: {volume domain: (cc, rhoc)} -> ``BC.fluid`` (the equivalent-fluid branch, FEM_3D.py:1367-1421)."""
    ref = load_reference()
    AP = ref.controlsair.AirProperties(c0=c0, rho0=rho0)
    AC = ref.controlsair.AlgControls(AP, freq_vec=np.asarray(freq, dtype=np.float64))
    BC = ref.BoundaryConditions.BC(AC, AP)
    for g, mu in mu_by_group.items():
        BC.mu[g] = np.asarray(mu)
    for g, v in (v_by_group or {}).items():
        BC.v[g] = np.asarray(v)
    for d, (cc, rhoc) in (fluid or {}).items():
        BC.fluid(d, cc, rhoc)
    S = PlainSource(src_coord, src_q)
    R = PlainReceiver(rcv_coord)
    obj = ref.FEM_3D.FEM3D(grid, S, R, AP, AC, BC)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        obj.compute(timeit=False)
        obj.evaluate(R)
    return obj
