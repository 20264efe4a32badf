"""The headless supervisor loop and the PheromoneAlgorithm drop-in (swarm_robotics_pheromones_b200/headless.py)
over the CUDA trail engine (TrailSwarm, csrc/pf_trail.cu): the 760-iteration replay of Phase-3/heatmap.txt must
reproduce the real reference's golden trace (tests/golden/sup_replay_heatmap.json). That the same drop-in and the
same world reproduce it under the UNMODIFIED Controller.main is checked where the reference checkout exists
(tests/test_headless_cpu.py)."""
import json

import pytest

from swarm_robotics_pheromones_b200 import headless
from swarm_robotics_pheromones_b200.trail import TrailSwarm
from tests.oracle_trail_engine import OracleTrailEngine

pytestmark = pytest.mark.gpu


def test_headless_loop_on_cuda_trail_engine_matches_reference_golden(golden_dir):
    g = json.load(open(golden_dir / "sup_replay_heatmap.json"))
    n = len(g["ticks"])
    world = headless.HeadlessWorld(food=g["food"], t_first=g["t0"], max_iterations=n)
    world.add_robot("e2", g["positions"][0], ps=g["ps"], stream=g["positions"])
    eng = TrailSwarm(["e2"], [g["start"]], g["food"], ps=g["ps"], start_delay=5.0)
    log = headless.run_supervisor_loop(world, eng, ["e2"], n)
    assert len(log) == n
    for entry, want in zip(log, g["ticks"]):
        rec = entry["records"]["e2"]
        assert entry["t"] == want["t"]
        assert rec["amount"] == want["amount"] and rec["n_global"] == want["n_global"]      # exact
        assert rec["direction"] == want["direction"] and rec["n_live"] == want["n_live"]    # exact
        assert {k: [e[0] for e in v] for k, v in rec["sense"].items()} == \
               {k: [e[0] for e in v] for k, v in want["sense"].items()}                      # list membership + order
        for a, b in zip(entry["speeds"]["e2"], want["speeds"]):
            assert abs(a - b) <= 1e-12
        assert abs(rec["sum_live"] - want["sum_live"]) <= 1e-12 * max(1.0, want["sum_live"])
    eng.close()


def test_dropin_class_with_default_cuda_engine_drives_three_robots():
    # three virtual e-pucks, integrated motion, one engine tick per iteration for all of them; the CUDA engine
    # against the CPU trail oracle behind the same loop
    def run(make_engine):
        world = headless.HeadlessWorld(food=(0.0, -0.59, 0.05), t_first=0.0, food_radius=0.08, arena=(-1, 1, -1, 1))
        starts = {"e1": (0.3, 0.2, 0.0), "e2": (-0.4, 0.1, 0.0), "e3": (0.02, -0.30, 0.0)}
        for nm, p in starts.items():
            world.add_robot(nm, p, theta=-1.5708 if nm == "e3" else 0.5, ps=60.0)
        eng = make_engine(list(starts), list(starts.values()), world.food)
        return headless.run_supervisor_loop(world, eng, list(starts), 150), eng

    cuda_log, eng = run(lambda names, starts, food: TrailSwarm(names, starts, food, ps=60.0))
    cpu_log, _ = run(lambda names, starts, food: OracleTrailEngine(names, starts, food))
    assert len(cuda_log) == len(cpu_log) == 150
    for a, b in zip(cuda_log, cpu_log):
        assert a["t"] == b["t"]
        for nm in ("e1", "e2", "e3"):
            for x, y in zip(a["speeds"][nm], b["speeds"][nm]):
                assert abs(x - y) <= 1e-9
            if b["records"]:
                ra, rb = a["records"][nm], b["records"][nm]
                assert ra["amount"] == rb["amount"] and ra["n_global"] == rb["n_global"]
                assert ra["direction"] == rb["direction"] and ra["n_live"] == rb["n_live"]
    eng.close()
