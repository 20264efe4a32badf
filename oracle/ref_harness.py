"""Headless driver for the UNMODIFIED reference controller (Phase-3/pheromone_algo_supervisor.py).

TEST INFRASTRUCTURE ONLY. It imports the reference from a checkout (default /root/reference, which
exists only in the build container, never on the GPU box) with two stub modules standing in for
Webots' `controller` and for `matplotlib`, and replays a position stream through
Controller.startPheromone / Controller.homing exactly as Controller.main does (sup:704-754).
It is used by tests/golden/make_golden.py to produce the committed golden traces and by
tests/test_reference_live.py (skipped when the checkout is absent).

Nothing from the reference is copied: the module is loaded from where it lies.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import tempfile
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("PF_REFERENCE_ROOT", "/root/reference"))
SUP_PATH = "Phase-3/pheromone_algo_supervisor.py"


def available() -> bool:
    return (REFERENCE_ROOT / SUP_PATH).exists()


def _install_stubs():
    if "controller" not in sys.modules:
        m = types.ModuleType("controller")
        for name in ("Motor", "Field", "DistanceSensor", "LED", "Node", "Camera", "Emitter",
                     "Receiver", "Supervisor", "Robot"):
            setattr(m, name, type(name, (), {}))
        sys.modules["controller"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        for fn in ("plot", "title", "xlabel", "ylabel", "grid", "savefig", "imread", "subplots"):
            setattr(plt, fn, lambda *a, **k: None)
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def load_module(relpath: str = SUP_PATH, name: str = "_ref_sup"):
    _install_stubs()
    spec = importlib.util.spec_from_file_location(name, str(REFERENCE_ROOT / relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Node:
    def __init__(self, pos):
        self.pos = list(pos)

    def getPosition(self):
        return list(self.pos)

    def getField(self, _name):
        return self

    def getSFVec3f(self):
        return list(self.pos)


class _Motor:
    def __init__(self):
        self.v = 0.0

    def setVelocity(self, v):
        self.v = v


class _Supervisor:
    def __init__(self, name, start, food):
        self.name = name
        self.node = _Node(start)
        self.food = _Node(food)
        self.t = 0.0

    def getSelf(self):
        return self.node

    def getFromDef(self, _d):
        return self.food

    def getName(self):
        return self.name

    def getTime(self):
        return self.t

    def getDevice(self, _n):
        return None


class FakeRobot:
    """The attributes of InitializeRobot (sup:11-178) that the pheromone path reads."""

    def __init__(self, name, start, food, ps):
        self.supervisor = _Supervisor(name, start, food)
        self.MAX_SPEED = 3.14  # sup:74
        self.coefficients = [[0.942, -0.22], [0.63, -0.1], [0.5, -0.06], [-0.06, -0.06],
                             [-0.06, -0.06], [-0.06, 0.5], [-0.19, 0.63], [-0.13, 0.942]]  # sup:77
        self.distance_sensors_values = [float(ps)] * 8
        self.speeds = [0.0, 0.0]
        self.left_motor = _Motor()
        self.right_motor = _Motor()


class ReferenceDriver:
    """One robot (`e2`) driven tick by tick through the real reference classes."""

    def __init__(self, start, food=(0.36, -0.15, 0.05), ps=60.0, name="e2"):
        self.mod = load_module()
        self.robot = FakeRobot(name, start, food, ps)
        self.mod.robot = self.robot  # module global read at sup:256,420,457
        self._cwd = tempfile.mkdtemp(prefix="pf_ref_")
        with self._quiet():
            self.algo = self.mod.PheromoneAlgorithm(self.robot)
            self.ctrl = self.mod.Controller(self.robot, self.algo)
            self.algo.setController(self.ctrl)
        self.mod.controller = self.ctrl  # module global read at sup:297,331,385
        self.name = name

    @contextlib.contextmanager
    def _quiet(self):
        old = os.getcwd()
        os.chdir(self._cwd)  # deposit_pheromone writes pheromone_grid.txt into the CWD (sup:293)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                yield
        finally:
            os.chdir(old)

    def set_state(self, homing: bool):
        self.ctrl.robot_state = self.mod.RobotState.HOMING if homing else self.mod.RobotState.SEARCHING

    def _install_recorders(self):
        """Wrap the bound methods ON THE INSTANCE so each tick records what the reference itself
        computed mid-tick (the sense dict is taken before that tick's evaporation)."""
        algo = self.algo
        rec = self._rec = {}
        dep, sen, adj = algo.deposit_pheromone, algo.sense_pheromone, algo.adjust_robot_behavior

        def deposit(name, pos):
            out = dep(name, pos)
            t = self.ctrl.time["current_time"]
            live = algo.pheromone_grids.get(name, {})
            rec["amount"] = live[t][0] if t in live and live[t][1] == pos else None
            return out

        def sense(name, pos):
            out = sen(name, pos)
            rec["sense"] = {k: [[it[0], it[1][0]] for it in v] for k, v in out.items()}
            best, direction = None, None
            for key, lst in out.items():  # the arg-max of sup:443-450, restated for the trace
                for it in lst:
                    if best is None or it[1][0] > best:
                        best, direction = it[1][0], key
            rec["direction"] = direction
            return out

        def adjust(sensed, name):
            out = adj(sensed, name)
            rec["speeds"] = [out[0], out[1]]
            return out

        algo.deposit_pheromone, algo.sense_pheromone, algo.adjust_robot_behavior = deposit, sense, adjust

    def tick(self, t, pos):
        """sup:711-754 without rendering: advance time, place the robot, run the pheromone path."""
        c = self.ctrl
        if not hasattr(self, "_rec"):
            self._install_recorders()
        self._rec.clear()
        self.robot.supervisor.t = t
        self.robot.supervisor.node.pos = list(pos)
        c.present_time = t
        c.time["previous_time"] = c.time["current_time"]
        c.time["current_time"] = t
        with self._quiet():
            if c.robot_state == self.mod.RobotState.SEARCHING:
                if t >= 5:
                    c.startPheromone(["e1", "e2"])
            else:
                c.homing(["e1", "e2"])
        glob = self.algo.global_pheromone_grid.get(self.name, {})
        live = self.algo.pheromone_grids.get(self.name, {})
        return {
            "t": t,
            "amount": self._rec.get("amount"),          # None = "nothing deposited"
            "n_global": len(glob),
            "sense": self._rec.get("sense"),
            "direction": self._rec.get("direction"),
            "speeds": self._rec.get("speeds"),
            "n_live": len(live),
            "sum_live": sum(v[0] for v in live.values()),
            "live": [[k, v[0]] for k, v in live.items()],
        }

