"""CUDA postprocessing (through the Python API -> C ABI) against the reference's known
answers and reference-produced golden vectors: bit-exact."""
from types import SimpleNamespace

import numpy as np
import pytest
from numpy import testing as nptest

import adacharge_b200 as ab
from adacharge_b200.interface import InfrastructureInfo
from oracle import postprocessing as opp
from tests.conftest import infra_from_json
from tests.test_oracle_postprocessing import REALLOC_KATS, realloc_kat_inputs

pytestmark = pytest.mark.gpu


def _infra(n, max_pilot=None, allowable=None):
    return InfrastructureInfo(np.zeros((0, 0)), np.zeros(0), np.zeros(n), np.full(n, 208.0), [], [str(i) for i in range(n)],
                              np.full(n, 32.0) if max_pilot is None else np.asarray(max_pilot, dtype=float), np.zeros(n),
                              allowable if allowable is not None else [np.array([0.0, 8, 16, 24, 32])] * n)


@pytest.mark.parametrize("value,expected", [(16, 16), (33, 32), (-1, 0)])
def test_project_continuous_kats(require_gpu, value, expected):  # t_pp.py:103-123
    out = ab.project_into_continuous_feasible_pilots(np.full((5, 20), value), _infra(5))
    nptest.assert_equal(out, expected)
    assert out.dtype == np.full((5, 20), value).dtype


@pytest.mark.parametrize("value,expected", [(16, 16), (18, 16), (15.98, 16), (33, 32), (-1, 0)])
def test_project_discrete_kats(require_gpu, value, expected):  # t_pp.py:126-157
    nptest.assert_equal(ab.project_into_discrete_feasible_pilots(np.full((5, 20), value), _infra(5)), expected)


@pytest.mark.parametrize("name", list(REALLOC_KATS))
def test_index_based_reallocation_kats(require_gpu, name):  # t_pp.py:173-318
    iface, rates, peak, expected = realloc_kat_inputs(name)
    out = ab.index_based_reallocation(rates, iface.active_sessions(), iface.infrastructure_info(), peak, ab.earliest_deadline_first, iface)
    nptest.assert_equal(out, expected)
    assert out is rates  # mutates its input like pp.py:183


def test_golden_projections_bit_exact(require_gpu, pp_golden):
    for c in pp_golden["project_continuous"]:
        r = np.array(c["rates"], dtype=c.get("dtype", "float64"))
        out = ab.project_into_continuous_feasible_pilots(r, _infra(r.shape[0], max_pilot=c["max_pilot"]))
        nptest.assert_array_equal(out, np.array(c["expected"]))
    for c in pp_golden["project_discrete"]:
        r = np.array(c["rates"], dtype=c.get("dtype", "float64"))
        out = ab.project_into_discrete_feasible_pilots(r, _infra(r.shape[0], allowable=[np.asarray(a, dtype=float) for a in c["allowable_pilots"]]))
        nptest.assert_array_equal(out, np.array(c["expected"]))


def test_golden_reallocation_bit_exact(require_gpu, pp_golden):
    for kind in ("diff_based", "index_based"):
        for c in pp_golden[kind]:
            iface = ab.TestingInterface({"active_sessions": c["active_sessions"], "infrastructure_info": infra_from_json(c["infrastructure_info"]),
                                         "current_time": 0, "period": 5})
            S, I = iface.active_sessions(), iface.infrastructure_info()
            r = np.array(c["rates"], dtype=float)
            if kind == "diff_based":
                out = ab.diff_based_reallocation(r, S, I, iface)
            else:
                out = ab.index_based_reallocation(r, S, I, c["peak_limit"], ab.earliest_deadline_first, iface)
            nptest.assert_array_equal(out, np.array(c["expected"]), err_msg=f"{kind} {c['network']}")


def test_golden_feasibility(require_gpu, pp_golden):
    for c in pp_golden["feasible"]:
        I = ab.TestingInterface({"infrastructure_info": infra_from_json(c["infrastructure_info"]), "active_sessions": [], "period": 5}).infrastructure_info()
        r = np.array(c["rates"], dtype=float)
        assert ab.infrastructure_constraints_feasible(r[:, 0], I) == c["expected_col0"]
        assert ab.infrastructure_constraints_feasible(r, I) == c["expected_all"]


def test_random_large_matches_oracle(require_gpu):
    """Full-size 54 x 288 schedule and the 1000-EVSE site: CUDA == numpy restatement, bit for bit."""
    from adacharge_b200.generators import config_c2, config_c5

    for cfg, seed in ((config_c2, 11), (config_c5, 12)):
        d = cfg(seed)
        iface = ab.TestingInterface(d)
        S, I = iface.active_sessions(), iface.infrastructure_info()
        rng = np.random.default_rng(seed)
        T = 24
        r = np.zeros((I.num_stations, T))
        for s in S:
            r[I.get_station_index(s.station_id)] = rng.uniform(0, 12, T)
        nptest.assert_array_equal(ab.project_into_discrete_feasible_pilots(r, I), opp.project_into_discrete_feasible_pilots(r, I))
        nptest.assert_array_equal(ab.project_into_continuous_feasible_pilots(r * 4 - 5, I), opp.project_into_continuous_feasible_pilots(r * 4 - 5, I))
        nptest.assert_array_equal(ab.diff_based_reallocation(r, S, I, iface), opp.diff_based_reallocation(r, S, I, iface))


def test_idempotence_and_membership(require_gpu):
    from adacharge_b200.generators import config_c2

    iface = ab.TestingInterface(config_c2(21))
    I = iface.infrastructure_info()
    r = np.random.default_rng(3).uniform(-2, 40, (54, 288))
    d1 = ab.project_into_discrete_feasible_pilots(r, I)
    nptest.assert_array_equal(ab.project_into_discrete_feasible_pilots(d1, I), d1)
    for i in range(54):
        assert np.isin(d1[i], I.allowable_pilots[i]).all()
    c1 = ab.project_into_continuous_feasible_pilots(r, I)
    nptest.assert_array_equal(ab.project_into_continuous_feasible_pilots(c1, I), c1)
    assert c1.min() >= 0 and c1.max() <= 32
