"""The headless Webots stand-in and the supervisor drop-in (swarm_robotics_pheromones_b200/headless.py) under the
UNMODIFIED reference: its own InitializeRobot() and Controller.main() (Phase-3/pheromone_algo_supervisor.py:11-178,
704-754) run on HeadlessWorld and must reproduce the golden 760-tick replay of Phase-3/heatmap.txt
(tests/golden/sup_replay_heatmap.json, made from the reference's startPheromone by make_golden.py) —
  1. with the reference's own PheromoneAlgorithm                 (pins the fake Webots API)
  2. with the drop-in PheromoneAlgorithm over a trail engine      (pins the drop-in's call protocol)
  3. and this package's statement of the loop, run_supervisor_loop, gives the same records.
  4. the multi-robot variants of the same file (Phase-3/experiments.py, Phase-3/experiment2.py; their loop is
     experiments.py:939-963), unmodified, replay the same golden trace on the headless world — with their own
     PheromoneAlgorithm and with the drop-in — and two controllers of experiments.py running in lockstep (one per
     robot, as Webots runs them) command exactly the wheel speeds run_supervisor_loop commands for two robots.
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


# ---- the multi-robot variants (Phase-3/experiments.py, Phase-3/experiment2.py: SURVEY §2 "cross-check only") --------
# Same classes, parametrised by robot name (InitializeRobot(name, supervisor), Controller(robot, pheromone, name),
# main() = ONE iteration, the `while supervisor.step(...)` loop lives in the file's __main__ block). The classes are
# wired here the way sup:757-762 wires them — one Controller object. (The variants' own __main__ blocks append a SECOND
# Controller to the list they loop over, experiments.py:947-951, while the pheromone object and the module-level
# `controller` that deposit_pheromone reads the clock from keep pointing at the first one; that driver code is
# Webots-only and out of scope.)
def _load_variant(world, relpath, name):
    saved = {k: sys.modules.get(k) for k in ("controller", "matplotlib", "matplotlib.pyplot")}
    ref_harness._install_stubs()
    sys.modules["controller"] = world.controller_module(name)
    try:
        return ref_harness.load_module(relpath, name=f"_ref_{relpath[-8:-3]}_{name}_{id(world)}")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _wire_variant(mod, world, name, make_algo):
    sup = world.controller_module(name).Supervisor()
    robot = mod.InitializeRobot(name=name, supervisor=sup)   # the variant's own device set-up, on the fake devices
    mod.robot = robot
    algo = make_algo(mod, robot)
    ctrl = mod.Controller(robot, algo, name)
    algo.setController(ctrl)
    mod.controller = ctrl
    return sup, robot, algo, ctrl


@needs_reference
@pytest.mark.parametrize("dropin", [False, True])
@pytest.mark.parametrize("relpath", ["Phase-3/experiments.py", "Phase-3/experiment2.py"])
def test_unmodified_multi_robot_variants_replay_the_golden_trace(golden, relpath, dropin):
    world = _world(golden, N_ITER)
    mod = _load_variant(world, relpath, "e2")
    records = []
    with _scratch():
        if dropin:
            eng = OracleTrailEngine(["e2"], [golden["start"]], golden["food"], ps=golden["ps"])
            make = lambda m, robot: headless.PheromoneAlgorithm(robot, engine=eng)   # noqa: E731
        else:
            make = lambda m, robot: m.PheromoneAlgorithm(robot)                      # noqa: E731
        sup, robot, algo, ctrl = _wire_variant(mod, world, "e2", make)

        def hook(name, it):
            live = algo.pheromone_grids.live(name) if hasattr(algo.pheromone_grids, "live") else \
                algo.pheromone_grids.get(name, {})
            records.append({"t": ctrl.time["current_time"], "speeds": [robot.left_motor.getVelocity(),
                                                                      robot.right_motor.getVelocity()],
                            "n_live": len(live), "sum_live": sum(v[0] for v in live.values())})

        world.on_iteration = hook
        try:
            while sup.step(robot.timestep) != -1:     # experiments.py:961-963
                ctrl.main()
        except SystemExit:
            pass
    _check(records, golden, N_ITER)


@needs_reference
def test_two_controllers_of_the_multi_robot_variant_in_lockstep_equal_the_package_loop():
    # two e-pucks, one controller "process" each (a thread with its own module instance, as Webots runs one process
    # per robot), the world steps when both have stepped; motion by the diff-drive stand-in, e2 reaches the food and
    # goes HOMING. The package's run_supervisor_loop on an identical world must command the same wheel speeds at
    # every iteration and leave the robots in the same place.
    import threading
    n_iter = 140
    starts = {"e1": (0.3, 0.2, 0.0), "e2": (0.02, -0.30, 0.0)}
    thetas = {"e1": 0.5, "e2": -1.5708}
    food = (0.0, -0.59, 0.05)

    def make_world():
        w = headless.HeadlessWorld(food=food, t_first=0.0, food_radius=0.08, arena=(-1, 1, -1, 1), max_iterations=n_iter)
        for n, p in starts.items():
            w.add_robot(n, p, theta=thetas[n], ps=60.0)
        return w

    world = make_world()
    world.lockstep(2)
    mods = {n: _load_variant(world, "Phase-3/experiments.py", n) for n in starts}
    log = {n: [] for n in starts}
    errors = []

    def run(n):
        try:
            sup, robot, algo, ctrl = _wire_variant(mods[n], world, n, lambda m, robot: m.PheromoneAlgorithm(robot))
            while sup.step(robot.timestep) != -1:
                ctrl.main()
                log[n].append((ctrl.time["current_time"], [robot.left_motor.getVelocity(), robot.right_motor.getVelocity()],
                               ctrl.robot_state))
        except SystemExit:
            pass
        except BaseException as e:                    # never leave the other thread waiting at the barrier
            errors.append(repr(e))
            world._barrier.abort()

    with _scratch():
        threads = [threading.Thread(target=run, args=(n,), daemon=True) for n in starts]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=120)
    assert not errors and all(len(log[n]) == n_iter for n in starts), (errors, {n: len(v) for n, v in log.items()})
    assert log["e2"][-1][2] == headless.HOMING and log["e1"][-1][2] == headless.SEARCHING

    world2 = make_world()
    eng = OracleTrailEngine(list(starts), list(starts.values()), food)
    plog = headless.run_supervisor_loop(world2, eng, list(starts), n_iter)
    assert len(plog) == n_iter
    for it, entry in enumerate(plog):
        for n in starts:
            t, speeds, _ = log[n][it]
            assert t == entry["t"] and speeds == entry["speeds"][n], (it, n, speeds, entry["speeds"][n])
    for n in starts:
        assert world.bodies[n].pos[:2] == world2.bodies[n].pos[:2]


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
