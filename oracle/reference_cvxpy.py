"""TEST INFRASTRUCTURE ONLY (nothing under adacharge_b200/ may import this).

The real reference solver path, when it can run: `AdaptiveChargingOptimization.solve` of the UNMODIFIED reference
(adacharge/adaptive_charging_optimization.py:286-321) loaded by path from `baseline/_ref/adacharge` or
`/root/reference/adacharge`, which needs cvxpy and a conic solver (reference setup.py:24, unpinned).  Neither is in this image
or its wheelhouse, so `available()` is False here and every caller falls back to the restatement in oracle/mpc.py; the hook
exists so that the oracle pin (tests/test_oracle_mpc.py) and the CPU arm of bench.py switch to the reference by themselves on a
machine where `import cvxpy` succeeds.  NOT EXERCISED in this image (parity at the solver boundary stays "unpinned", see the
header of oracle/mpc.py).

acnportal is used by the reference file for three type names only (aco.py:5); when it is missing, a module of that name
exporting this repo's stand-in classes is registered first — the same shim tests/golden/make_golden.py uses for the
reference's postprocessing.  Nothing of the reference is copied."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = (os.path.join(ROOT, "baseline", "_ref", "adacharge"), "/root/reference/adacharge")
_mod = None


def _ref_dir():
    for d in REF_DIRS:
        if os.path.exists(os.path.join(d, "adaptive_charging_optimization.py")):
            return d
    return None


def available() -> bool:
    """True when the reference's own solve can run here: its source is present and cvxpy imports."""
    if _ref_dir() is None:
        return False
    try:
        import cvxpy  # noqa: F401
    except Exception:
        return False
    return True


def load():
    """The reference's adaptive_charging_optimization module (cached)."""
    global _mod
    if _mod is not None:
        return _mod
    d = _ref_dir()
    if d is None:
        raise RuntimeError("reference source not found")
    try:
        import acnportal.acnsim.interface  # noqa: F401
    except Exception:
        from adacharge_b200 import interface as shim

        for name in ("acnportal", "acnportal.acnsim", "acnportal.acnsim.interface"):
            sys.modules.setdefault(name, types.ModuleType(name))
        m = sys.modules["acnportal.acnsim.interface"]
        m.Interface, m.SessionInfo, m.InfrastructureInfo = shim.Interface, shim.SessionInfo, shim.InfrastructureInfo
    spec = importlib.util.spec_from_file_location("adacharge_ref_aco", os.path.join(d, "adaptive_charging_optimization.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["adacharge_ref_aco"] = mod
    spec.loader.exec_module(mod)
    _mod = mod
    return mod


def solve_reference(objective, sessions, infra, interface, constraint_type="SOC", enforce_energy_equality=False,
                    peak_limit=None, prev_peak=0, solver=None):
    """`objective` in the oracle's form [(name, coefficient, kwargs), ...] (names of the reference's objective functions,
    aco.py:336-408).  Returns the reference's (N, T) schedule; raises its InfeasibilityException."""
    ref = load()
    comps = [ref.ObjectiveComponent(getattr(ref, name), coef, dict(kw)) for name, coef, kw in objective]
    kwargs = {} if solver is None else {"solver": solver}
    aco = ref.AdaptiveChargingOptimization(comps, interface, constraint_type, enforce_energy_equality, **kwargs)
    return aco.solve(sessions, infra, peak_limit=peak_limit, prev_peak=prev_peak)
