"""Pins the postprocessing oracle against every known answer of the reference's
tests/test_postprocessing.py and against golden vectors produced by the reference's own
postprocessing.py (tests/golden/make_golden.py).  CPU only."""
from types import SimpleNamespace

import numpy as np
import pytest
from numpy import testing as nptest

from adacharge_b200.generators import session_generator, single_phase_single_constraint, three_phase_balanced_network
from adacharge_b200.interface import TestingInterface, earliest_deadline_first
from oracle import postprocessing as opp
from tests.conftest import infra_from_json

SET = np.array([0, 5, 10])


@pytest.mark.parametrize("x,eps,expected", [(5, 0.05, 5), (5, 0, 5), (4.9, 0.05, 0), (4.98, 0.05, 5), (-1, 0.05, 0), (15, 0.05, 10), (4.95, 0.05, 0)])
def test_floor_to_set(x, eps, expected):  # t_pp.py:17-46
    assert opp.floor_to_set(x, SET, eps=eps) == expected


@pytest.mark.parametrize("x,eps,expected", [(5, 0.05, 5), (5, 0, 5), (2.5, 0.05, 5), (5.02, 0.05, 5), (-1, 0.05, 0), (15, 0.05, 10)])
def test_ceil_to_set(x, eps, expected):  # t_pp.py:49-78
    assert opp.ceil_to_set(x, SET, eps=eps) == expected


@pytest.mark.parametrize("x,expected", [(5, 10), (2.5, 5), (-1, 0), (15, 10)])
def test_increment_in_set(x, expected):  # t_pp.py:81-100
    assert opp.increment_in_set(x, SET) == expected


def _mock_infra():
    return SimpleNamespace(max_pilot=np.full(5, 32), min_pilot=np.full(5, 0), allowable_pilots=[[0, 8, 16, 24, 32]] * 5, num_stations=5)


@pytest.mark.parametrize("value,expected", [(16, 16), (33, 32), (-1, 0)])
def test_project_continuous_kats(value, expected):  # t_pp.py:103-123
    nptest.assert_equal(opp.project_into_continuous_feasible_pilots(np.full((5, 20), value), _mock_infra()), expected)


@pytest.mark.parametrize("value,expected", [(16, 16), (18, 16), (15.98, 16), (33, 32), (-1, 0)])
def test_project_discrete_kats(value, expected):  # t_pp.py:126-157
    nptest.assert_equal(opp.project_into_discrete_feasible_pilots(np.full((5, 20), value), _mock_infra()), expected)


REALLOC_KATS = {
    # name: (infra builder, remaining_energy, peak, expected first column)   t_pp.py:173-318
    "peak_binding": (lambda: single_phase_single_constraint(3, 66, allowable_pilots=[np.array([0, 8, 16, 24, 32])] * 3), [3.3] * 3, 48, [16, 16, 16]),
    "infra_not_binding": (lambda: single_phase_single_constraint(3, 66, allowable_pilots=[np.array([0] + list(range(8, 33)))] * 3), [3.3] * 3, 50, [17, 17, 16]),
    "single_phase_binding": (lambda: single_phase_single_constraint(3, 49, allowable_pilots=[np.array([0] + list(range(8, 33)))] * 3), [3.3] * 3, 60, [17, 16, 16]),
    "three_phase_binding": (lambda: three_phase_balanced_network(1, 16.51 * np.sqrt(3), allowable_pilots=[np.array([0] + list(range(8, 33)))] * 3), [3.3] * 3, 60, [17, 16, 16]),
    "energy_binding": (lambda: single_phase_single_constraint(3, 66, allowable_pilots=[np.array([0] + list(range(8, 33)))] * 3), [0.277, 3.3, 3.3], 50, [16, 17, 17]),
}


def realloc_kat_inputs(name):
    mk, remaining, peak, col0 = REALLOC_KATS[name]
    sessions = session_generator(3, [0] * 3, [2, 3, 4], [3.3] * 3, remaining, [32] * 3, min_rates=[0] * 3)
    iface = TestingInterface({"active_sessions": sessions, "infrastructure_info": mk(), "current_time": 0, "period": 5})
    expected = np.full((3, 10), 16)
    expected[:, 0] = col0
    return iface, np.full((3, 10), 16), peak, expected


@pytest.mark.parametrize("name", list(REALLOC_KATS))
def test_index_based_reallocation_kats(name):
    iface, rates, peak, expected = realloc_kat_inputs(name)
    out = opp.index_based_reallocation(rates, iface.active_sessions(), iface.infrastructure_info(), peak, earliest_deadline_first, iface)
    nptest.assert_equal(out, expected)


# ---------------------------------------------------------------- reference-produced vectors
def test_golden_projections(pp_golden):
    for c in pp_golden["project_continuous"]:
        r = np.array(c["rates"], dtype=c.get("dtype", "float64"))
        infra = SimpleNamespace(max_pilot=np.array(c["max_pilot"]), num_stations=r.shape[0])
        out = opp.project_into_continuous_feasible_pilots(r, infra)
        nptest.assert_array_equal(out, np.array(c["expected"]))
    for c in pp_golden["project_discrete"]:
        r = np.array(c["rates"], dtype=c.get("dtype", "float64"))
        infra = SimpleNamespace(allowable_pilots=c["allowable_pilots"], num_stations=r.shape[0])
        out = opp.project_into_discrete_feasible_pilots(r, infra)
        nptest.assert_array_equal(out, np.array(c["expected"]))


def test_golden_reallocation(pp_golden):
    for kind in ("diff_based", "index_based"):
        for c in pp_golden[kind]:
            iface = TestingInterface({"active_sessions": c["active_sessions"], "infrastructure_info": infra_from_json(c["infrastructure_info"]),
                                      "current_time": 0, "period": 5})
            S, I = iface.active_sessions(), iface.infrastructure_info()
            r = np.array(c["rates"], dtype=float)
            if kind == "diff_based":
                out = opp.diff_based_reallocation(r, S, I, iface)
            else:
                out = opp.index_based_reallocation(r, S, I, c["peak_limit"], earliest_deadline_first, iface)
            nptest.assert_array_equal(out, np.array(c["expected"]))


def test_golden_feasibility(pp_golden):
    for c in pp_golden["feasible"]:
        d = infra_from_json(c["infrastructure_info"])
        I = SimpleNamespace(phases=d["phases"], constraint_matrix=d["constraint_matrix"], constraint_limits=d["constraint_limits"])
        r = np.array(c["rates"], dtype=float)
        assert opp.infrastructure_constraints_feasible(r[:, 0], I) == c["expected_col0"]
        assert opp.infrastructure_constraints_feasible(r, I) == c["expected_all"]
