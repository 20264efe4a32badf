"""The GPU sparse-trail engine (pf_trail_*) against the REAL reference's golden traces: every
discrete quantity exact (amounts, counts, sense-list membership and order, arg-max direction,
decision), real-valued ones to 1e-12."""
import json

import numpy as np
import pytest

from oracle import trail as trail_oracle
from swarm_robotics_pheromones_b200.trail import TrailSwarm

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-12):
    return abs(a - b) <= tol * max(1.0, abs(b))


def _check(rec, want, where):
    assert rec["amount"] == want["amount"], where
    assert rec["n_global"] == want["n_global"], where
    assert rec["n_live"] == want["n_live"], where
    assert rec["direction"] == want["direction"], where
    assert _close(rec["sum_live"], want["sum_live"]), where
    if want["speeds"] is None:
        return
    assert _close(rec["speeds"][0], want["speeds"][0]) and _close(rec["speeds"][1], want["speeds"][1]), where
    for key in ("current", "front", "back", "left", "right"):
        got, exp = rec["sense"][key], want["sense"][key]
        assert len(got) == len(exp), (where, key)
        for g, e in zip(got, exp):
            assert g[0] == e[0] and _close(g[1], e[1]), (where, key)


def test_searching_replay_matches_reference(golden_dir):
    g = json.load(open(golden_dir / "sup_replay_heatmap.json"))
    sw = TrailSwarm(["e2"], [g["start"]], g["food"], ps=g["ps"], start_delay=5.0)
    for rec, pos in zip(g["ticks"], g["positions"]):
        out = sw.startPheromone(["e1", "e2"], {"e2": pos}, t_now=rec["t"])["e2"]
        _check(out, rec, f"t={rec['t']}")
        assert out["decision"] == "STRAIGHT"
    final = sw.pheromone_grid("e2")
    want = g["ticks"][-1]["live"]
    assert [k for k in final] == [w[0] for w in want]
    sw.close()


def test_homing_replay_matches_reference(golden_dir):
    g = json.load(open(golden_dir / "sup_replay_homing.json"))
    sw = TrailSwarm(["e2"], [g["start"]], g["food"], ps=g["ps"], states=2, start_delay=5.0)
    o = trail_oracle.TrailOracle(g["start"], g["food"], ps=(g["ps"],) * 8)
    o.state = trail_oracle.HOMING
    seen = set()
    for rec, pos in zip(g["ticks"], g["positions"]):
        out = sw.homing(["e2"], {"e2": pos}, t_now=rec["t"])["e2"]
        _check(out, rec, f"t={rec['t']}")
        want_dec = o.tick(rec["t"], pos)["decision"]
        assert out["decision"] == want_dec
        seen.add(want_dec)
    assert seen == {"STRAIGHT", "LEFT", "RIGHT", "KEEP", "STOP"}
    sw.close()


def test_many_robots_same_trace_and_start_delay(golden_dir):
    """N robots replaying the same trajectory with per-robot offsets must each reproduce the
    single-robot record; SEARCHING robots are inert before the 5 s start delay (sup:731-735)."""
    g = json.load(open(golden_dir / "sup_replay_heatmap.json"))
    n = 257
    names = [f"r{i}" for i in range(n)]
    sw = TrailSwarm(names, [g["start"]] * n, g["food"], ps=g["ps"], start_delay=5.0)
    pos0 = np.tile(np.array(g["positions"][0]), (n, 1))
    r = sw.startPheromone(None, pos0, t_now=4.9)
    assert all(v["amount"] is None and v["n_live"] == 0 and v["direction"] is None for v in r.values())
    for rec, pos in list(zip(g["ticks"], g["positions"]))[:150]:
        out = sw.startPheromone(None, np.tile(np.array(pos), (n, 1)), t_now=rec["t"])
        for name in (names[0], names[100], names[-1]):
            _check(out[name], rec, f"{name} t={rec['t']}")
    sw.close()


def test_full_live_list_is_reported_not_silent():
    # a deposit that finds the robot's live list full used to be dropped silently. With a capacity of
    # 8 entries the ninth deposit (entries only start to evaporate after 2 s) raises PF_ESTATE through TrailSwarm
    from swarm_robotics_pheromones_b200 import abi
    from swarm_robotics_pheromones_b200._ffi import PFError
    sw = TrailSwarm(["e2"], [[0.0, 0.0, 0.0]], [0.36, -0.15, 0.05], ps=60.0, start_delay=0.0, capacity=8)
    t, y = 0.0, 0.0
    with pytest.raises(PFError) as e:
        for _ in range(12):
            y += 0.004
            sw.startPheromone(["e2"], {"e2": [0.0, y, 0.0]}, t_now=t)
            t += 0.064
    assert e.value.code == abi.PF_ESTATE and "live list" in str(e.value)
    sw.close()
