"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol of
include/adacharge_b200.h, objective / session packing, API error rules, fixtures."""
import ctypes
import os
import re

import numpy as np
import pytest

import adacharge_b200 as ab
from adacharge_b200 import _cabi, engine
from adacharge_b200.generators import (
    session_generator, single_phase_single_constraint, three_phase_balanced_network, caltech_acn_infrastructure,
    hierarchical_three_phase_network, config_c1, config_c2, config_c5,
)
from oracle import mpc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "adacharge_b200.h")).read()
    declared = set(re.findall(r"\b(acb_[a-z_]+)\s*\(", hdr))
    assert declared >= set(_cabi.EXPORTS)
    L = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _cabi.lib().acb_version() >= 100


def test_default_options_struct_layout():
    o = _cabi.default_options(max_iter=123, equality=1)
    assert o.max_iter == 123 and o.equality == 1 and abs(o.eps_rel - 1e-4) < 1e-9
    with pytest.raises(TypeError):
        _cabi.default_options(bogus=1)


def test_pack_objective_matches_oracle_terms():
    d = config_c2(3)
    iface = ab.TestingInterface(d)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    T = mpc.horizon(S)
    ext = np.linspace(0, 40, T)
    obj = [ab.ObjectiveComponent(ab.quick_charge, 2.0), ab.ObjectiveComponent(ab.equal_share, 0.1),
           ab.ObjectiveComponent(ab.tou_energy_cost, 1.5), ab.ObjectiveComponent(ab.total_energy, 3.0),
           ab.ObjectiveComponent(ab.demand_charge, 0.7), ab.ObjectiveComponent(ab.load_flattening, 0.2, {"external_signal": ext}),
           ab.ObjectiveComponent(ab.non_completion_penalty, 0.4)]
    p = ab.pack_objective(obj, I, iface, T, prev_peak=iface.get_prev_peak())
    terms = mpc.objective_terms([(c.function.__name__, c.coefficient, c.kwargs) for c in obj], I, iface, T, S, iface.get_prev_peak())
    k = np.asarray(I.voltages) / 1e3
    lin = p["alpha"][None, :] + k[:, None] * p["beta"][None, :]
    # non_completion_penalty(norm=1) is linear: the oracle keeps it separate, add it here
    lin_oracle = terms["lin"] - 0.4 * (k * 5 / 60)[:, None]
    assert np.allclose(lin, lin_oracle, rtol=1e-12, atol=1e-12)
    assert p["qd"] == pytest.approx(terms["diag_q"])
    assert p["gamma"] == pytest.approx(0.2) and np.allclose(p["ext"], ext)
    assert p["peak_w"] == pytest.approx(0.7 * 15.51) and p["peak_p0"] == pytest.approx(terms["peaks"][0][1])


def test_numeric_objectives_equal_oracle_evaluation():
    d = config_c2(4)
    iface = ab.TestingInterface(d)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    T = mpc.horizon(S)
    R = np.random.default_rng(1).uniform(0, 32, (54, T))
    for fn in (ab.quick_charge, ab.equal_share, ab.tou_energy_cost, ab.total_energy, ab.demand_charge, ab.load_flattening,
               ab.non_completion_penalty):
        assert fn(R, I, iface) == pytest.approx(mpc.evaluate_objective(R, [(fn.__name__, 1, {})], I, iface, S), rel=1e-12)


def test_unknown_objective_is_rejected_loudly():
    iface = ab.TestingInterface(config_c1(0))
    I = iface.infrastructure_info()
    with pytest.raises(TypeError, match="not callable"):
        ab.pack_objective([ab.ObjectiveComponent("quick_charge")], I, iface, 10)
    # numeric callables are traced (tracer.py); what the packed objective cannot hold is rejected, not approximated
    for bad in (lambda rates, infra, iface, **kw: -float(np.abs(np.asarray(rates) - 3.0).sum()),          # not quadratic
                lambda rates, infra, iface, **kw: -float((np.asarray(rates)[:, 1:] * np.asarray(rates)[:, :-1]).sum()),  # couples periods
                lambda rates, infra, iface, **kw: -float((np.arange(np.shape(rates)[0])[:, None] * np.asarray(rates) ** 2).sum()),  # per-EVSE weights
                lambda rates, infra, iface, **kw: float((np.asarray(rates) ** 2).sum()),                     # convex
                lambda rates, **kw: 0.0,                                                                      # wrong signature
                lambda rates, infra, iface, **kw: rates.value):                                               # cvxpy-style
        with pytest.raises(TypeError, match="objective component"):
            ab.pack_objective([ab.ObjectiveComponent(bad)], I, iface, 10)


def test_traced_components_equal_the_builtin_specs():
    """tracer.py recovers the packed form of a numeric callable: every built-in, traced as if it were a user function,
    gives the spec the hand-written `_spec_*` gives (mixed-voltage site so alpha and beta separate)."""
    from adacharge_b200 import tracer
    from adacharge_b200.adaptive_charging_optimization import _SPECS
    d = config_c2(3)
    iface = ab.TestingInterface(d)
    I = iface.infrastructure_info()
    I.voltages = np.asarray(I.voltages, dtype=float).copy()
    I.voltages[::3] = 240.0
    T = 24
    ext = np.linspace(0, 50, T)
    for fn, kw in ((ab.quick_charge, {}), (ab.equal_share, {}), (ab.tou_energy_cost, {}), (ab.total_energy, {}),
                   (ab.load_flattening, {"external_signal": ext})):
        want = _SPECS[fn](I, iface, T, **kw)
        got = tracer.trace_component(lambda r, i, f, _fn=fn, **k: _fn(r, i, f, **k), I, iface, T, **kw)
        k = np.asarray(I.voltages) / 1e3
        # compare through the objective both describe (load_flattening's signal is folded into beta by the tracer)
        def lin(sp):
            b = np.asarray(sp.get("beta", np.zeros(T)), dtype=float) + 2 * sp.get("gamma", 0.0) * np.asarray(sp.get("ext", np.zeros(T)))
            return np.asarray(sp.get("alpha", np.zeros(T)), dtype=float)[None, :] + k[:, None] * b[None, :]
        np.testing.assert_allclose(lin(got), lin(want), rtol=1e-9, atol=1e-9)
        assert got.get("qd", 0.0) == pytest.approx(want.get("qd", 0.0), abs=1e-9)
        assert got.get("gamma", 0.0) == pytest.approx(want.get("gamma", 0.0), abs=1e-9)


def test_traced_user_component_packs_like_its_handwritten_twin():
    d = config_c2(1)
    iface = ab.TestingInterface(d)
    I = iface.infrastructure_info()
    T = 36
    w = np.linspace(1.0, 0.2, T)

    def solar_following(rates, infrastructure, interface, solar=None, **kw):
        u = (np.asarray(rates) * (np.asarray(infrastructure.voltages)[:, None] / 1e3)).sum(axis=0)
        return -float(((u - solar[: len(u)]) ** 2).sum()) + float(w[: len(u)] @ np.asarray(rates).sum(axis=0)) - 0.01 * float((np.asarray(rates) ** 2).sum())

    solar = 30 * np.sin(np.linspace(0, np.pi, T)) ** 2
    got = ab.pack_objective([ab.ObjectiveComponent(solar_following, 0.5, {"solar": solar})], I, iface, T)
    want = ab.pack_objective([ab.ObjectiveComponent(ab.load_flattening, 0.5, {"external_signal": -solar}),
                              ab.ObjectiveComponent(ab.equal_share, 0.005)], I, iface, T)
    want["alpha"] = want["alpha"] - 0.5 * w
    k = np.asarray(I.voltages) / 1e3
    lin = lambda o: o["alpha"][None, :] + k[:, None] * (o["beta"] + 2 * o["gamma"] * (o["ext"] if o["ext"] is not None else 0))[None, :]
    np.testing.assert_allclose(lin(got), lin(want), rtol=1e-9, atol=1e-9)
    assert got["qd"] == pytest.approx(want["qd"], abs=1e-10) and got["gamma"] == pytest.approx(want["gamma"], abs=1e-10)


def test_nonconcave_components_are_rejected():
    iface = ab.TestingInterface(config_c2(0))
    I = iface.infrastructure_info()
    with pytest.raises(ValueError):
        ab.pack_objective([ab.ObjectiveComponent(ab.peak, 1.0)], I, iface, 10)
    with pytest.raises(ValueError):
        ab.pack_objective([ab.ObjectiveComponent(ab.equal_share, -1.0)], I, iface, 10)


def test_component_kwargs_override_caller_kwargs():  # aco.py:203-217
    iface = ab.TestingInterface(config_c2(0))
    I = iface.infrastructure_info()
    p = ab.pack_objective([ab.ObjectiveComponent(ab.demand_charge, 1.0, {"baseline_peak": 500.0})], I, iface, 10, baseline_peak=1.0)
    assert p["peak_p0"] == 500.0


def test_pack_sessions_sorted_by_row_and_units():
    d = config_c2(7)
    iface = ab.TestingInterface(d)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    ps = engine.pack_sessions(S, I, iface.period)
    assert list(ps["sess_row"]) == sorted(ps["sess_row"])
    j = ps["order"][0]
    assert ps["sess_energy"][0] == pytest.approx(S[j].remaining_demand / (208 * 5 / 1e3 / 60))


def test_session_info_semantics():
    s = ab.SessionInfo("a", "b", 10, 4, arrival=3, departure=20, current_time=5, max_rates=32)
    assert (s.arrival_offset, s.remaining_time, s.remaining_demand) == (0, 15, 6)
    assert len(s.max_rates) == 15
    s = ab.SessionInfo("a", "b", 10, 4, arrival=8, departure=20, current_time=5)
    assert (s.arrival_offset, s.remaining_time) == (3, 12)


def test_solve_without_sessions_returns_zero_column():  # aco.py:310-311 (no GPU needed)
    iface = ab.TestingInterface(config_c1(0))
    out = ab.AdaptiveChargingOptimization([ab.ObjectiveComponent(ab.quick_charge)], iface).solve([], iface.infrastructure_info())
    assert out.shape == (30, 1) and not out.any()


def test_schedule_without_sessions_and_ctor_rules():  # ada.py:101-113, 137-138
    alg = ab.AdaptiveSchedulingAlgorithm([ab.ObjectiveComponent(ab.quick_charge)])
    assert alg.schedule([]) == {}
    with pytest.raises(ValueError):
        ab.AdaptiveSchedulingAlgorithm([], quantize=False, reallocate=True)
    with pytest.warns(UserWarning):
        a = ab.AdaptiveSchedulingAlgorithm([], quantize=True, max_recompute=5)
    assert a.max_recompute == 1


def test_offline_schedule_lookup_and_errors():  # ada.py:278-294, t_int.py:290-308
    from unittest.mock import Mock

    alg = ab.AdaptiveChargingAlgorithmOffline([])
    with pytest.raises(ValueError):
        alg.schedule([])
    alg.internal_schedule = {"s1": np.arange(10.0), "s2": np.arange(10.0) * 2}
    alg.session_ids = {"a", "b"}
    iface = Mock(); iface.current_time = 3
    alg.register_interface(iface)
    evs = [Mock(station_id="s1", session_id="a"), Mock(station_id="s2", session_id="b")]
    assert alg.schedule(evs) == {"s1": [3.0], "s2": [6.0]}
    with pytest.raises(ValueError):
        alg.schedule([Mock(station_id="s1", session_id="zzz")])


def test_fixture_networks():
    infra = three_phase_balanced_network(1, 16.51 * np.sqrt(3))
    a = mpc.soc_rows(ab.TestingInterface({"infrastructure_info": infra, "active_sessions": [], "period": 5}).infrastructure_info())
    cur = lambda r: np.hypot(a[:, 0] @ r, a[:, 1] @ r).max()
    assert cur(np.array([16.0, 16, 16])) == pytest.approx(27.7128, abs=1e-3)
    assert cur(np.array([17.0, 16, 16])) == pytest.approx(28.583, abs=1e-3)
    assert cur(np.array([17.0, 17, 16])) > 16.51 * np.sqrt(3)
    c = caltech_acn_infrastructure()
    assert c["constraint_matrix"].shape == (8, 54) and c["constraint_limits"][0] == pytest.approx(416.667, abs=1e-3)
    assert c["constraint_limits"][3] == pytest.approx(180.505, abs=1e-3)
    h = hierarchical_three_phase_network(1000)
    assert h["constraint_matrix"].shape == (50 + 30 + 6, 1000)


def test_generators_are_seeded_and_shaped():
    for cfg, T in ((config_c1, 144), (config_c2, 288), (config_c5, 288)):
        a, b = cfg(3), cfg(3)
        assert a["active_sessions"] == b["active_sessions"]
        iface = ab.TestingInterface(a)
        assert mpc.horizon(iface.active_sessions()) == T


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.NativeLibraryMissing):
        _cabi.lib()


@pytest.mark.parametrize("with_quick_charge", [True, False])
def test_fleet_replay_packing_matches_object_packing(with_quick_charge):
    """replay_fast packs with array operations what replay.SiteReplay packs through SessionInfo /
    build_instance: same session tables, energies, horizons and objective vectors."""
    import adacharge_b200 as ab
    from adacharge_b200.adaptive_charging_optimization import AdaptiveChargingOptimization
    from adacharge_b200.generators import caltech_acn_infrastructure
    from adacharge_b200.interface import TestingInterface
    from adacharge_b200.replay import SiteReplay
    from adacharge_b200.replay_fast import FleetReplay

    # quick_charge's coefficients depend on the horizon (one evaluation per distinct T); without it one evaluation serves all sites
    obj = [ab.ObjectiveComponent(ab.tou_energy_cost, 2.0), ab.ObjectiveComponent(ab.total_energy, 0.3), ab.ObjectiveComponent(ab.demand_charge, 1 / 30)]
    if with_quick_charge:
        obj.insert(0, ab.ObjectiveComponent(ab.quick_charge))
    infra = caltech_acn_infrastructure()
    slow = SiteReplay(infra, obj, n_sites=4, steps=288, seed0=50)
    fast = FleetReplay(infra, obj, n_sites=4, steps_per_day=288, seed0=50, Tp=160)
    fast.prev_peak[:] = slow.prev_peak[:] = [0.0, 120.0, 40.0, 300.0]
    for t in (20, 90, 150):
        h, idx, s, pos, n_sess = fast._pack(t)
        for b in range(4):
            sess = slow._active(b, t)
            assert len(sess) == n_sess[b]
            if not sess:
                assert h["T"][b] == 1
                continue
            iface = TestingInterface({"active_sessions": [], "infrastructure_info": infra, "current_time": t, "period": 5,
                                      "prices": slow.prices, "demand_charge": slow.demand_charge, "prev_peak": float(slow.prev_peak[b])})
            inst = AdaptiveChargingOptimization(obj, iface).build_instance(sess, slow.info, None, float(slow.prev_peak[b]))
            k = len(sess)
            assert h["T"][b] == inst.T
            np.testing.assert_array_equal(h["sess_row"][b, :k], inst.sess_row)
            np.testing.assert_array_equal(h["sess_len"][b, :k], inst.sess_len)
            np.testing.assert_array_equal(h["sess_start"][b, :k], inst.sess_start)
            np.testing.assert_allclose(h["sess_energy"][b, :k], inst.sess_energy.astype(np.float32), rtol=1e-7)
            np.testing.assert_allclose(h["alpha"][b, : inst.T], inst.alpha.astype(np.float32), rtol=1e-7)
            np.testing.assert_allclose(h["beta"][b, : inst.T], inst.beta.astype(np.float32), rtol=1e-7)
            assert (h["alpha"][b, inst.T:] == 0).all() and (h["beta"][b, inst.T:] == 0).all()
            assert h["peak_w"][b] == np.float32(inst.peak_w) and h["peak_p0"][b] == np.float32(inst.peak_p0)
            ro = h["sess_rate_off"][b, :k]
            assert (ro < 0).all()
            np.testing.assert_array_equal(h["max_rates"][-(ro + 1)], [mx[0] for mx in inst.max_rates])
            assert all((mx == mx[0]).all() for mx in inst.max_rates)


def test_custom_objective_component_protocol():
    """User-defined components carry a `kernel_spec` (the restatement of the reference's open-ended
    ObjectiveComponent.function, aco.py:200-218); anything else is rejected loudly; kwargs merge like aco.py:203-217."""
    import adacharge_b200 as ab
    from adacharge_b200.adaptive_charging_optimization import pack_objective
    from tests.scenarios import SCENARIOS, make_interface

    sc = SCENARIOS["tiny_feasible"]
    iface = make_interface(sc)
    I = iface.infrastructure_info()

    def late_charge(rates, infrastructure, interface, weight=1.0, **kw):
        T = np.shape(rates)[1]
        return float(weight * (np.arange(T) / T) @ np.asarray(rates).sum(axis=0))

    late_charge.kernel_spec = lambda infra, interface, T, weight=1.0, **kw: dict(alpha=-weight * np.arange(T) / T)
    ob = pack_objective([ab.ObjectiveComponent(late_charge, 2.0, {"weight": 3.0}), ab.ObjectiveComponent(ab.equal_share, 0.5)], I, iface, 12, weight=100.0)
    np.testing.assert_allclose(ob["alpha"], -6.0 * np.arange(12) / 12)  # component kwargs win over caller kwargs
    assert ob["qd"] == 0.5
    with pytest.raises(TypeError, match="cannot be evaluated"):
        pack_objective([ab.ObjectiveComponent(lambda rates, **kw: 0.0)], I, iface, 12)


def test_fleet_replay_shards_are_the_unsharded_fleet():
    """tools/replay_c4.py gives rank r the sites shard_range(n, r, world) through `site_offset`: the shards'
    EV tables concatenate to the single-process fleet (same seeds per site), so sharding changes no result."""
    import adacharge_b200 as ab
    from adacharge_b200.generators import caltech_acn_infrastructure
    from adacharge_b200.replay_fast import FleetReplay
    from adacharge_b200.sharding import shard_range

    obj = [ab.ObjectiveComponent(ab.quick_charge)]
    infra = caltech_acn_infrastructure()
    full = FleetReplay(infra, obj, n_sites=7, days=2, seed0=40, Tp=128)
    parts = []
    for r in range(3):
        rg = shard_range(7, r, 3)
        parts.append((rg, FleetReplay(infra, obj, n_sites=len(rg), days=2, seed0=40, Tp=128, site_offset=rg.start)))
    assert sum(len(rg) for rg, _ in parts) == 7
    # tables are day-major: compare after sorting both sides by (site, arrival)
    site_cat = np.concatenate([p.ev_site + rg.start for rg, p in parts])
    arr_cat = np.concatenate([p.ev_arr for _, p in parts])
    oc, of = np.lexsort((arr_cat, site_cat)), np.lexsort((full.ev_arr, full.ev_site))
    np.testing.assert_array_equal(site_cat[oc], full.ev_site[of])
    for name in ("ev_station", "ev_arr", "ev_dep", "ev_req", "ev_max"):
        np.testing.assert_array_equal(np.concatenate([getattr(p, name) for _, p in parts])[oc], getattr(full, name)[of])
    # and the packed batch of a step is the concatenation of the shards' batches
    h_full = full._pack(130)[0]
    hs = [p._pack(130)[0] for _, p in parts]
    for k in ("T", "n_sessions", "sess_row", "sess_len", "sess_energy", "alpha", "beta", "peak_w", "peak_p0"):
        np.testing.assert_array_equal(np.concatenate([h[k] for h in hs]), h_full[k])


def test_status_mapping_and_inaccurate_exit():
    """aco.py:319-320: anything but optimal / optimal_inaccurate raises.  Iteration-limit exits with a small
    certified gap and violation count as 'inaccurate' and are returned."""
    from adacharge_b200.adaptive_charging_optimization import check_status

    check_status(_cabi.ACB_SOLVED, dict(gap=1.0, violation=1.0))
    with pytest.warns(RuntimeWarning, match="iteration limit"):
        check_status(_cabi.ACB_MAX_ITER, dict(gap=5e-3, violation=5e-6))
    with pytest.raises(ab.InfeasibilityException):  # 5e-4 is above the 1e-5 parity bar ...
        check_status(_cabi.ACB_MAX_ITER, dict(gap=5e-3, violation=5e-4))
    with pytest.warns(RuntimeWarning):               # ... unless the caller opts in
        check_status(_cabi.ACB_MAX_ITER, dict(gap=5e-3, violation=5e-4), violation=1e-3)
    for status, info in ((_cabi.ACB_MAX_ITER, dict(gap=5e-2, violation=0.0)), (_cabi.ACB_MAX_ITER, dict(gap=0.0, violation=1e-2)),
                         (_cabi.ACB_INFEASIBLE, dict(gap=0.0, violation=0.0)), (_cabi.ACB_NUMERICAL, dict(gap=0.0, violation=0.0))):
        with pytest.raises(ab.InfeasibilityException, match="Solve failed with status"):
            check_status(status, info)


def test_options_struct_matches_header_order():
    """The ctypes mirror of acb_options lists the header's fields in the header's order (a silent mismatch would
    shift every option after it)."""
    hdr = open(os.path.join(ROOT, "include", "adacharge_b200.h")).read()
    body = hdr[hdr.index("typedef struct acb_options {"):hdr.index("} acb_options;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:float|int32_t)\s+(\w+)\s*;", body)
    assert fields == [n for n, _ in _cabi.Options._fields_]
    body = hdr[hdr.index("typedef struct acb_batch {"):hdr.index("} acb_batch;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in re.findall(r"(?:const\s+)?(?:int32_t|float|double)\s*\*?\s*([^;]+);", body):
        names += [x.strip().lstrip("*").split()[-1].lstrip("*") for x in decl.split(",")]
    assert names == [n for n, _ in _cabi.Batch._fields_], (names, [n for n, _ in _cabi.Batch._fields_])
    d = _cabi.default_options()
    assert d.eps_rel == pytest.approx(1e-4) and d.check_every == 25 and d.term_floor == pytest.approx(0.05) and d.rho_curv == pytest.approx(1.0)
    assert d.rate_tol == pytest.approx(3e-4) and d.polish_min_qd == pytest.approx(5e-4)
    # the device packer's structs (acb_sessions, acb_objective) against the header as well
    for cname, ctype in (("acb_sessions", _cabi.Sessions), ("acb_objective", _cabi.Objective), ("acb_fleet", _cabi.Fleet)):
        body = hdr[hdr.index("typedef struct %s {" % cname):hdr.index("} %s;" % cname)]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in re.findall(r"(?:const\s+)?(?:int32_t|float|double)\s*\*?\s*([^;]+);", body):
            names += [re.sub(r"\[.*\]", "", x.strip().lstrip("*").split()[-1].lstrip("*")) for x in decl.split(",")]
        assert names == [n for n, _ in ctype._fields_], (cname, names)


def test_padded_horizon_and_long_horizons():
    """Horizons up to 288 use the on-chip kernel's instantiations, longer ones the next multiple of 32 (general path);
    the reference puts no limit on T (aco.py:243-245), this library's is 3616 periods."""
    from adacharge_b200 import engine

    assert [engine.padded_horizon(t) for t in (1, 64, 65, 128, 129, 160, 161, 288)] == [64, 64, 128, 128, 160, 160, 288, 288]
    assert engine.padded_horizon(289) == 320 and engine.padded_horizon(2016) == 2016 and engine.padded_horizon(3600) == 3616
    with pytest.raises(ValueError, match="longest supported horizon"):
        engine.padded_horizon(3617)


def test_quadratic_non_completion_penalty_packs_per_session_weights():
    """non_completion_penalty(norm=2) -> one weight per session: coefficient * (kWh per A*period of its EVSE)^2, in the
    packed (row-sorted) session order; the numeric twin and the oracle evaluate the same value."""
    d = config_c2(4)
    iface = ab.TestingInterface(d)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    spec = [("tou_energy_cost", 1.0, {}), ("non_completion_penalty", 2.5, {"norm": 2})]
    aco = ab.AdaptiveChargingOptimization([ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in spec], iface)
    inst = aco.build_instance(S, I)
    w = np.asarray(I.voltages)[inst.sess_row] * iface.period / 1e3 / 60
    np.testing.assert_allclose(inst.sess_quad, 2.5 * w * w)
    rng = np.random.default_rng(0)
    R = rng.uniform(0, 8, size=(I.num_stations, mpc.horizon(S)))
    val = sum(c * getattr(ab, n)(R, I, iface, **k) for n, c, k in spec)
    assert val == pytest.approx(mpc.evaluate_objective(R, spec, I, iface, S), rel=1e-12)
    with pytest.raises(ValueError, match="not concave"):
        ab.pack_objective([ab.ObjectiveComponent(ab.non_completion_penalty, -1.0, {"norm": 2})], I, iface, 10)


def test_peak_terms_with_different_baselines_are_kept_as_pieces():
    iface = ab.TestingInterface(config_c2(0))
    I = iface.infrastructure_info()
    ob = ab.pack_objective([ab.ObjectiveComponent(ab.demand_charge, 0.5, {"baseline_peak": 30.0}), ab.ObjectiveComponent(ab.peak, -2.0, {"baseline_peak": 20.0}),
                            ab.ObjectiveComponent(ab.demand_charge, 0.25, {"baseline_peak": 30.0})], I, iface, 10)
    dc = iface.get_demand_charge()
    assert ob["peak_terms"] == [(20.0, pytest.approx(2.0)), (30.0, pytest.approx(0.75 * dc))]   # (the site's previous peak is 17.1 kW)
    assert ob["peak_w"] == pytest.approx(2.0 + 0.75 * dc) and ob["peak_p0"] == 30.0   # the top piece
    one = ab.pack_objective([ab.ObjectiveComponent(ab.demand_charge, 0.5, {"baseline_peak": 30.0}), ab.ObjectiveComponent(ab.peak, -2.0, {"baseline_peak": 30.0})], I, iface, 10)
    assert "peak_terms" not in one and one["peak_p0"] == 30.0


def test_batched_objective_components_validation():
    from adacharge_b200.batched import objective_components

    comps = objective_components([ab.ObjectiveComponent(ab.quick_charge, 2.0), ab.ObjectiveComponent(ab.demand_charge, 0.5, {"baseline_peak": 7.0}),
                                  ab.ObjectiveComponent(ab.non_completion_penalty, 1.0, {"norm": 2})])
    assert comps == [(_cabi.OBJ_KIND["quick_charge"], 2.0, 0.0), (_cabi.OBJ_KIND["demand_charge"], 0.5, 7.0), (_cabi.OBJ_KIND["non_completion_penalty_l2"], 1.0, 0.0)]
    with pytest.raises(TypeError, match="cannot be packed on the device"):
        objective_components([ab.ObjectiveComponent(lambda rates, infra, iface, **kw: 0.0)])
    with pytest.raises(NotImplementedError, match="different baselines"):
        objective_components([ab.ObjectiveComponent(ab.demand_charge, 1.0, {"baseline_peak": 3.0}), ab.ObjectiveComponent(ab.peak, -1.0, {"baseline_peak": 4.0})])


def test_set_rounding_helpers_keep_the_reference_contract():
    """floor_to_set / ceil_to_set / increment_in_set (reference postprocessing.py:10-74), incl. the documented corner:
    a member exactly eps above x is NOT reached by floor_to_set."""
    from oracle import postprocessing as opp

    s = [0, 5, 10]
    assert ab.floor_to_set(4.95, s) == 0 and ab.floor_to_set(4.96, s) == 5 and ab.floor_to_set(5, s) == 5 and ab.floor_to_set(-1, s) == 0 and ab.floor_to_set(11, s) == 10
    assert ab.ceil_to_set(5.04, s) == 5 and ab.ceil_to_set(5.06, s) == 10 and ab.ceil_to_set(20, s) == 10 and ab.ceil_to_set(-3, s) == 0
    assert ab.increment_in_set(5, s) == 10 and ab.increment_in_set(4.9, s) == 5 and ab.increment_in_set(10, s) == 10 and ab.increment_in_set(-1, s) == 0
    rng = np.random.default_rng(1)
    sets = [np.array([0.0, 6.0, 8.0, 16.0, 32.0]), np.arange(0, 33, 1.0), np.array([8.0])]
    for a in sets:
        for x in np.concatenate([rng.uniform(-2, 36, 200), a, a + 0.05, a - 0.05]):
            assert ab.floor_to_set(x, a) == opp.floor_to_set(x, a)
            assert ab.ceil_to_set(x, a) == opp.ceil_to_set(x, a)
            assert ab.increment_in_set(x, a) == opp.increment_in_set(x, a)
