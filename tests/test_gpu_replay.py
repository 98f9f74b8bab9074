"""Closed-loop replay with warm starts (BASELINE config 4 shape, small): every step solves,
warm starts cut the iteration count, and the EVs get their energy."""
import numpy as np
import pytest

import adacharge_b200 as ab
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.replay import SiteReplay

pytestmark = pytest.mark.gpu


def _run(warm, objective):
    rp = SiteReplay(caltech_acn_infrastructure(), objective, n_sites=6, steps=288, seed0=100, warm_start=warm)
    stats = rp.run(96, 136)
    return rp, stats


def test_replay_warm_start_reduces_iterations(require_gpu):
    obj = [ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3), ab.ObjectiveComponent(ab.demand_charge, 1 / 30)]
    _, cold = _run(False, obj)
    _, warm = _run(True, obj)
    for st in cold.status + warm.status:
        assert (st == 0).all()
    ic = np.concatenate(cold.iters).mean()
    iw = np.concatenate(warm.iters).mean()
    print(f"mean iterations per step: cold {ic:.0f}, warm {iw:.0f}")
    assert iw < 0.8 * ic


def test_replay_quick_charge_delivers_energy(require_gpu):
    obj = [ab.ObjectiveComponent(ab.quick_charge), ab.ObjectiveComponent(ab.equal_share, 1e-6)]
    rp = SiteReplay(caltech_acn_infrastructure(), obj, n_sites=4, steps=288, seed0=7, warm_start=True)
    stats = rp.run(60, 288)
    for st in stats.status:
        assert (st == 0).all()
    assert (stats.delivered_frac >= 0.9999).all(), stats.delivered_frac  # t_int.py:37-39 asserts >= 99.99 %
