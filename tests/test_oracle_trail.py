"""The sparse-trail oracle (oracle/trail.py) pinned against the REAL reference: golden traces made by
tests/golden/make_golden.py from the unmodified Phase-3/pheromone_algo_supervisor.py."""
import json
import math

import pytest

from oracle import ref_harness, trail


def _replay(golden, homing):
    o = trail.TrailOracle(golden["start"], golden["food"], ps=(golden["ps"],) * 8)
    if homing:
        o.state = trail.HOMING
    decisions = []
    for rec, pos in zip(golden["ticks"], golden["positions"]):
        r = o.tick(rec["t"], pos)
        for k in ("amount", "n_global", "sense", "direction", "speeds", "n_live", "sum_live"):
            assert r[k] == rec[k], f"t={rec['t']}: {k}: oracle {r[k]} != reference {rec[k]}"
        if "live" in rec:
            assert r["live"] == rec["live"]
        decisions.append(r.get("decision"))
    return decisions


def test_searching_replay_of_shipped_trajectory(golden_dir):
    g = json.load(open(golden_dir / "sup_replay_heatmap.json"))
    assert len(g["ticks"]) == 760
    dec = _replay(g, homing=False)
    assert set(dec) == {"STRAIGHT"}  # the report's "no change needed" (PDF p.54)
    amounts = [t["amount"] for t in g["ticks"]]
    assert amounts.count(None) == 84          # 675/759 steps >= 1.5 mm + the first tick
    assert amounts.count(20.0) == 225 and amounts.count(1.0) == 451  # every-3rd rule, sup:316-318
    # wheel speeds at ps_j = 60
    assert g["ticks"][100]["speeds"] == [2.38778125, 2.3877812499999997]


def test_homing_replay_covers_every_branch(golden_dir):
    g = json.load(open(golden_dir / "sup_replay_homing.json"))
    dec = _replay(g, homing=True)
    assert {"STRAIGHT", "LEFT", "RIGHT", "KEEP", "STOP"} <= set(dec)
    assert all(t["amount"] in (None, 5.0) for t in g["ticks"])  # homing intensity, sup:346-348


def test_pheromonelevels_known_answers(golden_dir):
    """Phase-3/pheromonelevels.txt: I0 * exp(-0.1 * sum of ages), compounded once per 0.064 s tick
    while 2 <= age <= 4 (sup:594-607)."""
    lv = json.load(open(golden_dir / "pheromonelevels.json"))["e2"]
    assert len(lv) == 36

    def compound(i0, nticks):
        v = i0
        for k in range(nticks):
            v = trail.evaporate(2.048 + 0.064 * k, v)
        return v

    for want, (i0, n) in zip(lv[:7], [(20.0, 14), (20.0, 13), (20.0, 12), (20.0, 11), (1.0, 9), (1.0, 8), (1.0, 7)]):
        assert abs(compound(i0, n) - want) < 1e-12
    assert set(lv[7:]) == {1.0, 20.0}
    # deletion thresholds: 10th / 16th / 20th window tick
    assert compound(1.0, 9) > 0.1 >= compound(1.0, 10)
    assert compound(5.0, 15) > 0.1 >= compound(5.0, 16)
    assert compound(20.0, 19) > 0.1 >= compound(20.0, 20)


def test_pid_matches_reference_formula():
    # e = wrap(atan2(dy,dx) - pi/2), out = wrap(1.01 e)  (sup:199-226)
    for cur, goal in [((0, 0), (1, 0)), ((0.3, -0.4), (0.36, -0.15)), ((1, 1), (-2, 0.5)), ((0, 0), (0, -1))]:
        a = math.atan2(goal[1] - cur[1], goal[0] - cur[0]) - math.pi / 2
        e = math.atan2(math.sin(a), math.cos(a))
        out = 1.01 * e
        out = math.atan2(math.sin(out), math.cos(out))
        assert abs(trail.pid_turn(cur, goal) - out) < 1e-15


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present (GPU box)")
def test_live_reference_agrees_with_golden(golden_dir):
    """Re-drive the unmodified reference for the first 120 ticks and compare with the fixture."""
    g = json.load(open(golden_dir / "sup_replay_heatmap.json"))
    drv = ref_harness.ReferenceDriver(g["start"], food=tuple(g["food"]), ps=g["ps"])
    for rec, pos in list(zip(g["ticks"], g["positions"]))[:120]:
        r = drv.tick(rec["t"], pos)
        for k in ("amount", "n_global", "sense", "direction", "speeds", "n_live", "sum_live"):
            assert r[k] == rec[k]
