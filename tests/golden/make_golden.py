"""Generates tests/golden/postprocessing_golden.json by running the REFERENCE's own
adacharge/postprocessing.py and adacharge/utils.py (loaded by path from
/root/reference, which exists only in the build container) on seeded inputs.

acnportal is absent, and the two reference files use it only for type names in
annotations (pp.py:6, utils.py:2), so a module named acnportal.acnsim.interface that
exports this repo's stand-in classes is registered before loading them.  Nothing from
the reference is copied: only its inputs/outputs are stored.

Also writes tests/golden/mpc_oracle_golden.json: ORACLE-produced (not reference-
produced; cvxpy/ECOS cannot run here) optimal objectives and schedules of a few
seeded instances, so that the GPU parity tests need not re-run the slow oracle.

Run:  python tests/golden/make_golden.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/adacharge"

from adacharge_b200 import interface as shim  # noqa: E402
from adacharge_b200.generators import (  # noqa: E402
    session_generator, single_phase_single_constraint, three_phase_balanced_network, caltech_acn_infrastructure,
    config_c1, config_c2,
)


def load_reference():
    for name in ("acnportal", "acnportal.acnsim", "acnportal.acnsim.interface"):
        sys.modules.setdefault(name, types.ModuleType(name))
    m = sys.modules["acnportal.acnsim.interface"]
    m.Interface, m.SessionInfo, m.InfrastructureInfo = shim.Interface, shim.SessionInfo, shim.InfrastructureInfo
    m.__all__ = ["Interface", "SessionInfo", "InfrastructureInfo"]
    pkg = types.ModuleType("adacharge_ref")
    pkg.__path__ = [REF]
    sys.modules["adacharge_ref"] = pkg
    mods = {}
    for name in ("utils", "postprocessing"):
        spec = importlib.util.spec_from_file_location(f"adacharge_ref.{name}", os.path.join(REF, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"adacharge_ref.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["postprocessing"], mods["utils"]


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return jsonable(o.tolist())
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.bool_,)):
        return bool(o)
    return o


def realloc_case(rng, infra, n_sessions, T, kind):
    n = len(infra["station_ids"])
    stations = rng.permutation(n)[:n_sessions]
    arr = np.where(rng.random(n_sessions) < 0.8, 0, rng.integers(1, 4, n_sessions))
    dep = arr + rng.integers(2, T + 1, n_sessions)
    rem = rng.uniform(0.05, 6.0, n_sessions)
    req = rem + rng.uniform(0, 3, n_sessions)
    sessions = session_generator(
        n_sessions, arr.tolist(), dep.tolist(), req.tolist(), rem.tolist(), rng.choice([16.0, 32.0, 24.5], n_sessions).tolist(),
        station_ids=[infra["station_ids"][i] for i in stations],
        estimated_departures=(dep + rng.integers(-1, 2, n_sessions)).tolist(),
    )
    rates = np.zeros((n, T))
    rates[stations] = rng.uniform(0, 20, (n_sessions, T))
    if kind == "grid":  # already on allowable values
        rates[stations] = rng.choice([0.0, 8.0, 16.0], (n_sessions, T))
    return sessions, rates


def main():
    pp, utils = load_reference()
    rng = np.random.default_rng(20261018)
    out = {"project_continuous": [], "project_discrete": [], "index_based": [], "diff_based": [], "feasible": []}

    class Obj:
        pass

    # ---- projections
    for case in range(6):
        n, T = int(rng.integers(1, 9)), int(rng.integers(1, 12))
        infra = Obj()
        infra.num_stations = n
        infra.max_pilot = rng.choice([16.0, 32.0, 40.0, 6.5], n)
        sets = []
        for i in range(n):
            kind = rng.integers(0, 3)
            if kind == 0:
                sets.append([0, 8, 16, 24, 32])
            elif kind == 1:
                sets.append([0.0] + [float(v) for v in range(6, 33)])
            else:
                sets.append(sorted(set(np.round(rng.uniform(0, 40, int(rng.integers(1, 6))), 2).tolist())))
        infra.allowable_pilots = sets
        rates = rng.uniform(-3, 45, (n, T))
        # values on / near set members and near the eps boundary
        for _ in range(n * T // 2 + 1):
            i, t = int(rng.integers(0, n)), int(rng.integers(0, T))
            v = float(rng.choice(sets[i]))
            rates[i, t] = v + float(rng.choice([0.0, -0.05, -0.049999, -0.050001, 0.05, 1e-12, -1e-12, -0.02]))
        out["project_continuous"].append(dict(max_pilot=infra.max_pilot, rates=rates,
                                              expected=pp.project_into_continuous_feasible_pilots(rates, infra)))
        out["project_discrete"].append(dict(allowable_pilots=sets, rates=rates,
                                            expected=pp.project_into_discrete_feasible_pilots(rates, infra)))
    # integer-dtype inputs keep their dtype in the reference
    infra = Obj(); infra.num_stations = 3; infra.max_pilot = np.array([32, 32, 16]); infra.allowable_pilots = [[0, 8, 16, 24, 32]] * 3
    ri = np.array([[33, -1, 16], [18, 8, 7], [40, 15, 17]])
    out["project_continuous"].append(dict(max_pilot=infra.max_pilot, rates=ri, dtype="int64",
                                          expected=pp.project_into_continuous_feasible_pilots(ri, infra)))
    out["project_discrete"].append(dict(allowable_pilots=infra.allowable_pilots, rates=ri, dtype="int64",
                                        expected=pp.project_into_discrete_feasible_pilots(ri, infra)))

    # ---- reallocation on three network shapes
    nets = [
        ("single", lambda: single_phase_single_constraint(6, float(rng.uniform(40, 120)))),
        ("three", lambda: three_phase_balanced_network(3, float(rng.uniform(30, 90)))),
        ("three_ragged", lambda: three_phase_balanced_network(
            2, float(rng.uniform(25, 70)),
            allowable_pilots=[np.array([0, 8, 16, 24, 32]) if i % 2 else np.array([0] + list(range(6, 33))) for i in range(6)])),
        ("caltech", lambda: caltech_acn_infrastructure(transformer_cap=float(rng.uniform(20, 150)))),
    ]
    for name, mk in nets:
        for rep in range(4):
            infra_d = mk()
            n = len(infra_d["station_ids"])
            T = int(rng.integers(2, 8))
            ns = int(rng.integers(1, min(n, 30) + 1))
            sessions, rates = realloc_case(rng, infra_d, ns, T, "cont" if rep % 2 == 0 else "grid")
            iface = shim.TestingInterface({"active_sessions": sessions, "infrastructure_info": infra_d, "current_time": 0, "period": 5})
            S, I = iface.active_sessions(), iface.infrastructure_info()
            base = dict(network=name, infrastructure_info=infra_d, active_sessions=sessions, rates=rates)
            exp = pp.diff_based_reallocation(rates.copy(), S, I, iface)
            out["diff_based"].append(dict(base, expected=exp))
            rounded = pp.project_into_discrete_feasible_pilots(rates, I)
            peak = float(rounded[:, 0].sum() + rng.uniform(0, 40))
            exp = pp.index_based_reallocation(rounded.copy(), S, I, peak, shim.earliest_deadline_first, iface)
            out["index_based"].append(dict(base, rates=rounded, peak_limit=peak, expected=exp))
            out["feasible"].append(dict(network=name, infrastructure_info=infra_d, rates=rates,
                                        expected_col0=bool(utils.infrastructure_constraints_feasible(rates[:, 0], I)),
                                        expected_all=bool(utils.infrastructure_constraints_feasible(rates, I))))
    with open(os.path.join(HERE, "postprocessing_golden.json"), "w") as f:
        json.dump(jsonable(out), f)
    print({k: len(v) for k, v in out.items()})

    # ---- oracle-produced MPC vectors
    from oracle import mpc

    obj1 = [("quick_charge", 1, {}), ("equal_share", 1e-3, {})]
    obj2 = [("tou_energy_cost", 1, {}), ("total_energy", 0.3, {}), ("demand_charge", 1 / 30, {})]
    obj1s = [("quick_charge", 1, {}), ("equal_share", 0.05, {})]  # strongly concave: well-conditioned unique optimum
    gold = []
    for cfg, seed, cap in (("c1", 0, None), ("c1", 1, None), ("c1s", 0, None), ("c1s", 2, None), ("c2", 1, 150), ("c2", 2, 40), ("c2", 3, 60)):
        d = config_c1(seed) if cfg.startswith("c1") else config_c2(seed, infra=caltech_acn_infrastructure(transformer_cap=cap))
        iface = shim.TestingInterface(d)
        S, I = iface.active_sessions(), iface.infrastructure_info()
        obj = obj1 if cfg == "c1" else (obj1s if cfg == "c1s" else obj2)
        # the unique-optimum cases are compared rate by rate (1e-3 A): tightest tolerances the float64 IPM reaches
        R = mpc.solve_mpc(obj, S, I, iface, prev_peak=iface.get_prev_peak(), tol_scale=1e-4 if cfg.startswith("c1") else 1.0)
        gold.append(dict(config=cfg, seed=seed, transformer_cap=cap, objective=obj,
                         oracle_objective=mpc.evaluate_objective(R, obj, I, iface, S, iface.get_prev_peak()), rates=R))
        print(cfg, seed, cap, gold[-1]["oracle_objective"])
    with open(os.path.join(HERE, "mpc_oracle_golden.json"), "w") as f:
        json.dump(jsonable(gold), f)


if __name__ == "__main__":
    main()
