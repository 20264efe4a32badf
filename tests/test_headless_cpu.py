"""The headless Webots stand-in and the supervisor drop-in (swarm_robotics_pheromones_b200/headless.py) under the
UNMODIFIED reference: its own InitializeRobot() and Controller.main() (Phase-3/pheromone_algo_supervisor.py:11-178,
704-754) run on HeadlessWorld and must reproduce the golden 760-tick replay of Phase-3/heatmap.txt
(tests/golden/sup_replay_heatmap.json, made from the reference's startPheromone by make_golden.py) —
  1. with the reference's own PheromoneAlgorithm                 (pins the fake Webots API)
  2. with the drop-in PheromoneAlgorithm over a trail engine      (pins the drop-in's call protocol)
  3. and this package's statement of the loop, run_supervisor_loop, gives the same records.
The reference checkout only exists in the build container: 1 and 2 are skipped elsewhere. The CUDA engine behind
the same drop-in and loop is checked on the GPU box (tests/test_gpu_headless.py)."""
import contextlib
import io
import json
import os
import sys
import tempfile

import pytest

from oracle import ref_harness
from swarm_robotics_pheromones_b200 import headless
from tests.oracle_trail_engine import OracleTrailEngine

needs_reference = pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")


@pytest.fixture()
def golden(golden_dir):
    return json.load(open(golden_dir / "sup_replay_heatmap.json"))


def _world(g, iterations):
    w = headless.HeadlessWorld(food=g["food"], t_first=g["t0"], max_iterations=iterations)
    w.add_robot("e2", g["positions"][0], ps=g["ps"], stream=g["positions"][:iterations])
    return w


def _load_reference(world, name="e2"):
    """the unmodified file, with the world's module standing where Webots' `controller` would"""
    saved = {k: sys.modules.get(k) for k in ("controller", "matplotlib", "matplotlib.pyplot")}
    ref_harness._install_stubs()                      # matplotlib only (rendering is out of scope)
    sys.modules["controller"] = world.controller_module(name)
    try:
        return ref_harness.load_module(name=f"_ref_sup_{id(world)}")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@contextlib.contextmanager
def _scratch():
    old = os.getcwd()
    with tempfile.TemporaryDirectory(prefix="pf_headless_") as d:
        os.chdir(d)                                   # the reference writes heatmap.txt / pheromone_grid.txt
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                yield
        finally:
            os.chdir(old)


def _run_real_main(mod, world, make_algo):
    """sup:757-762 with the pheromone object of the caller's choice; returns per-iteration records"""
    records = []
    with _scratch():
        robot = mod.InitializeRobot()                 # the reference's own device set-up, on the fake devices
        mod.robot = robot
        algo = make_algo(mod, robot)
        ctrl = mod.Controller(robot, algo)
        algo.setController(ctrl)
        mod.controller = ctrl

        def hook(name, it):
            live = algo.pheromone_grids.live(name) if hasattr(algo.pheromone_grids, "live") else \
                algo.pheromone_grids.get(name, {})
            records.append({"t": ctrl.time["current_time"], "speeds": [robot.left_motor.getVelocity(),
                                                                      robot.right_motor.getVelocity()],
                            "n_live": len(live), "sum_live": sum(v[0] for v in live.values())})

        world.on_iteration = hook
        try:
            ctrl.main()                               # returns when the `while` step reports -1 (sup:705) ...
        except SystemExit:                            # ... or exits from robot.step() (sup:81-84)
            pass
    return records


def _check(records, g, n):
    assert len(records) == n
    for rec, want in zip(records, g["ticks"][:n]):
        assert rec["t"] == want["t"]
        assert rec["speeds"] == want["speeds"], (rec, want["speeds"])   # set_actuators re-sends robot.speeds (sup:753)
        assert rec["n_live"] == want["n_live"]
        assert abs(rec["sum_live"] - want["sum_live"]) <= 1e-12 * max(1.0, want["sum_live"])


N_ITER = 760


@needs_reference
def test_unmodified_main_with_reference_algorithm_on_headless_world(golden):
    world = _world(golden, N_ITER)
    mod = _load_reference(world)
    records = _run_real_main(mod, world, lambda m, robot: m.PheromoneAlgorithm(robot))
    _check(records, golden, N_ITER)


@needs_reference
def test_unmodified_main_with_dropin_algorithm(golden):
    world = _world(golden, N_ITER)
    mod = _load_reference(world)
    eng = OracleTrailEngine(["e2"], [golden["start"]], golden["food"], ps=golden["ps"])
    records = _run_real_main(mod, world, lambda m, robot: headless.PheromoneAlgorithm(robot, engine=eng))
    _check(records, golden, N_ITER)


def test_package_loop_matches_golden(golden):
    world = _world(golden, N_ITER)
    eng = OracleTrailEngine(["e2"], [golden["start"]], golden["food"], ps=golden["ps"])
    log = headless.run_supervisor_loop(world, eng, ["e2"], N_ITER)
    assert len(log) == N_ITER
    for entry, want in zip(log, golden["ticks"]):
        rec = entry["records"]["e2"]
        assert entry["t"] == want["t"] and entry["speeds"]["e2"] == want["speeds"]
        assert rec["amount"] == want["amount"] and rec["n_global"] == want["n_global"]
        assert rec["direction"] == want["direction"] and rec["n_live"] == want["n_live"]


def test_three_robots_integrated_motion_food_and_homing():
    # Phase-3/foraging_world.wbt-like: three e-pucks driven by the diff-drive stand-in; one heads for the food,
    # sees it through the camera stand-in and switches to HOMING (sup:609-620, 672-678)
    world = headless.HeadlessWorld(food=(0.0, -0.59, 0.05), t_first=0.0, food_radius=0.08, arena=(-1, 1, -1, 1))
    starts = {"e1": (0.3, 0.2, 0.0), "e2": (-0.4, 0.1, 0.0), "e3": (0.02, -0.30, 0.0)}
    for n, p in starts.items():
        world.add_robot(n, p, theta=-1.5708 if n == "e3" else 0.5, ps=60.0)
    eng = OracleTrailEngine(list(starts), list(starts.values()), world.food)
    seen = {"homing": False}

    def on_tick(it, t, res):
        seen["homing"] |= eng.o["e3"].state == 2
    log = headless.run_supervisor_loop(world, eng, list(starts), 120, on_tick=on_tick)
    assert len(log) == 120 and seen["homing"]
    assert log[-1]["t"] == 119 * 0.064
    assert log[119]["records"]["e3"]["amount"] in (5.0, None)       # HOMING deposits 5.0 (sup:346-348)
    moved = {n: abs(world.bodies[n].pos[0] - starts[n][0]) + abs(world.bodies[n].pos[1] - starts[n][1]) for n in starts}
    assert all(v > 0.05 for v in moved.values())      # 120 iterations at about 3 mm each
    assert log[100]["records"]["e1"]["n_global"] > 0 and log[10]["records"] == {}   # nothing before t = 5 s
