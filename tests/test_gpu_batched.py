"""Batched public API (adacharge_b200.batched.BatchedAdaptiveCharging): the device packer against the host packer
(bit-exact), the batched schedule against the drop-in class per instance, and the phased (two-launch) solve against
one launch.  Reference path: adacharge/adacharge.py:135-193 (schedule), adaptive_charging_optimization.py:200-321."""
import ctypes as C

import numpy as np
import pytest
import torch

import adacharge_b200 as ab
from adacharge_b200 import _cabi, engine
from adacharge_b200.batched import BatchedAdaptiveCharging, sessions_to_arrays
from adacharge_b200.generators import config_c2, caltech_acn_infrastructure

pytestmark = pytest.mark.gpu

BENCH_OBJ = [("tou_energy_cost", 1.0, {}), ("total_energy", 0.3, {}), ("demand_charge", 1.0 / 30.0, {})]
RICH_OBJ = [("quick_charge", 0.7, {}), ("equal_share", 0.02, {}), ("tou_energy_cost", 1.5, {}), ("total_energy", 0.3, {}),
            ("demand_charge", 1.0 / 30.0, {"baseline_peak": 12.0}), ("peak", -0.01, {"baseline_peak": 12.0}), ("load_flattening", 1e-3, {}),
            ("non_completion_penalty", 0.05, {})]


def _components(spec):
    return [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in spec]


def _cases(n, seed0=0):
    infra = caltech_acn_infrastructure()
    ifaces = [ab.TestingInterface(config_c2(seed0 + i, infra=infra, price_noise=0.2)) for i in range(n)]
    return infra, ifaces


def _batched_inputs(ifaces, Tp):
    I = ifaces[0].infrastructure_info()
    sess = sessions_to_arrays([f.active_sessions() for f in ifaces], I, S_max=54)
    prices = np.stack([np.asarray(f.get_prices(Tp), dtype=float) for f in ifaces])
    prev = np.array([f.get_prev_peak() for f in ifaces])
    return I, sess, prices, prev


@pytest.mark.parametrize("spec", [BENCH_OBJ, RICH_OBJ], ids=["bench", "every_component"])
def test_device_pack_equals_host_pack(require_gpu, spec):
    infra, ifaces = _cases(10)
    obj = _components(spec)
    I, sess, prices, prev = _batched_inputs(ifaces, 288)
    ext = 40.0 + 10.0 * np.sin(np.arange(288) / 17.0)
    bac = BatchedAdaptiveCharging(obj, I, 5, batch=len(ifaces), max_sessions=54, horizon=288, demand_charge=15.51, chunks=1)
    ch = bac.chunks[0]
    bac.upload_raw(sess, prices=prices, prev_peak=prev, external_signal=ext if bac.need_ext else None)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check(_cabi.lib().acb_pack_sessions(bac.site.handle, C.byref(ch.sessions), C.byref(ch.objective), C.byref(ch.batch), engine._ptr(ch.flags), st), "pack")
    torch.cuda.synchronize()
    assert int(ch.flags.cpu()[0]) == 0
    dev = {k: v.cpu().numpy() for k, v in ch.packed.items()}
    for b, iface in enumerate(ifaces):
        S = iface.active_sessions()
        comps = _components([(n, c, dict(k, **({"external_signal": ext} if n == "load_flattening" else {}))) for n, c, k in spec])
        aco = ab.AdaptiveChargingOptimization(comps, iface)
        inst = aco.build_instance(S, I, None, iface.get_prev_peak())
        pb = engine.PackedBatch(bac.site, [inst], Tp=288, S_max=54)
        h = {k: v.numpy() for k, v in pb.host.items()}
        n = len(S)
        assert dev["T"][b] == h["T"][0] and dev["n_sessions"][b] == n
        for k in ("sess_row", "sess_start", "sess_len", "sess_energy"):
            np.testing.assert_array_equal(dev[k][b, :n], h[k][0, :n], err_msg=k)
        for k in ("min_rates", "max_rates"):
            np.testing.assert_array_equal(dev[k][-dev["sess_rate_off"][b, :n] - 1], h[k][-h["sess_rate_off"][0, :n] - 1], err_msg=k)
        for k in ("alpha", "beta") + (("ext",) if bac.need_ext else ()):
            np.testing.assert_array_equal(dev[k][b], h[k][0], err_msg=k)
        for k in ("qd", "gamma", "peak_w", "peak_p0"):
            assert dev[k][b] == h[k][0], (k, dev[k][b], h[k][0])


def test_batched_schedule_equals_dropin_class_per_instance(require_gpu):
    infra, ifaces = _cases(9, seed0=50)
    obj = _components(BENCH_OBJ)
    I, sess, prices, prev = _batched_inputs(ifaces, 288)
    bac = BatchedAdaptiveCharging(obj, I, 5, batch=len(ifaces), max_sessions=54, horizon=288, demand_charge=15.51, chunks=3)
    res = bac.schedule(sess, prices=prices, prev_peak=prev)
    assert (res.status == 0).all()
    for b, iface in enumerate(ifaces):
        S = iface.active_sessions()
        aco = ab.AdaptiveChargingOptimization(obj, iface, solver_options=dict(eps_rel=1e-4))
        R = aco.solve(S, I, prev_peak=iface.get_prev_peak())
        pil = np.maximum(ab.project_into_continuous_feasible_pilots(R, I), 0)  # ada.py:177-190
        T = R.shape[1]
        assert res.T[b] == T
        np.testing.assert_array_equal(res.pilots[b, :, :T], pil)
        assert not res.pilots[b, :, T:].any()
    # reusable, deterministic
    again = bac.schedule(sess, prices=prices, prev_peak=prev).pilots.copy()
    np.testing.assert_array_equal(again, res.pilots)


def test_batched_flags_undeclared_multi_session(require_gpu):
    infra, ifaces = _cases(2, seed0=7)
    I, sess, prices, prev = _batched_inputs(ifaces, 288)
    sess["station"][1, 1] = sess["station"][1, 0]  # two sessions on one EVSE
    bac = BatchedAdaptiveCharging(_components(BENCH_OBJ), I, 5, batch=2, max_sessions=54, horizon=288, demand_charge=15.51, chunks=1)
    with pytest.raises(ValueError, match="multi_session"):
        bac.schedule(sess, prices=prices, prev_peak=prev)


def test_phased_solve_is_identical_to_one_launch(require_gpu):
    """acb_options.phase_iters parks every unfinished instance after the first launch and resumes it in a second one
    (ordered by remaining gap): same schedules, iteration counts and statuses, bit for bit."""
    infra, ifaces = _cases(200, seed0=300)
    obj = _components(BENCH_OBJ)
    insts = []
    for f in ifaces:
        aco = ab.AdaptiveChargingOptimization(obj, f)
        insts.append(aco.build_instance(f.active_sessions(), f.infrastructure_info(), None, f.get_prev_peak()))
    site = aco._site_for(ifaces[0].infrastructure_info(), insts[0])
    pb = engine.PackedBatch(site, insts).upload()
    pb.solve(_cabi.default_options(phase_iters=0))
    ref = [t.cpu().numpy().copy() for t in (pb.rates, pb.iters, pb.status, pb.stats)]
    assert (ref[2] == 0).all() and ref[1].max() > 150
    for K in (60, 100, 175):
        pb.rates.zero_()
        pb.solve(_cabi.default_options(phase_iters=K))
        for a, t in zip(ref, (pb.rates, pb.iters, pb.status, pb.stats)):
            np.testing.assert_array_equal(a, t.cpu().numpy(), err_msg=f"phase_iters={K}")
