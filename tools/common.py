"""Shared by the development scripts: the C3 benchmark instances through the host packer (engine.PackedBatch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BENCH_OBJECTIVE = [("tou_energy_cost", 1.0, {}), ("total_energy", 0.3, {}), ("demand_charge", 1.0 / 30.0, {})]


def build_instances(batch, seed0=0):
    import adacharge_b200 as ab
    from adacharge_b200.generators import config_c2, caltech_acn_infrastructure

    infra = caltech_acn_infrastructure()
    obj = [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in BENCH_OBJECTIVE]
    insts, site, ifaces = [], None, []
    for i in range(batch):
        iface = ab.TestingInterface(config_c2(seed0 + i, infra=infra, price_noise=0.2))
        S, I = iface.active_sessions(), iface.infrastructure_info()
        aco = ab.AdaptiveChargingOptimization(obj, iface)
        inst = aco.build_instance(S, I, None, iface.get_prev_peak())
        insts.append(inst)
        if site is None:
            site = aco._site_for(I, inst)
        if i < 64:
            ifaces.append(iface)
    return site, insts, ifaces
