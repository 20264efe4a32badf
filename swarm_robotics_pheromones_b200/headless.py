"""Headless stand-in for Webots + the supervisor-level drop-in.

Three pieces, so that the reference's controller loop can run without Webots against the B200 engines:

  HeadlessWorld          the part of Webots' `controller` API that Phase-3/pheromone_algo_supervisor.py (`sup`)
                         touches — Supervisor (step / getTime / getSelf / getFromDef / getDevice / getName), motors,
                         proximity sensors, LEDs, camera, displays — for N virtual e-pucks on one clock. Poses come
                         from a recorded stream (e.g. Phase-3/heatmap.txt) or from a diff-drive integration of the
                         motor velocities (E-puck.proto:86,100; the stand-in for Webots physics).
                         `world.controller_module(name)` is a module object to put at sys.modules['controller']
                         before the UNMODIFIED reference file is loaded: its InitializeRobot() (sup:11-178) and
                         Controller.main() (sup:704-754) then run as they are.
  PheromoneAlgorithm     drop-in for the reference class of that name (sup:228-525): same constructor argument, same
                         five methods and attributes Controller reads (get_robot_position, setController,
                         pheromone_grids, start_position ...). The arithmetic runs in a trail engine — TrailSwarm
                         (csrc/pf_trail.cu, all robots in one kernel) by default. The reference-side change is one line:
                             pheromone = PheromoneAlgorithm(robot)          # sup:759, now imported from here
  run_supervisor_loop    this package's own statement of Controller.main (sup:704-754) for N robots in one
                         process, for boxes where the reference checkout is not present; tests/test_headless_cpu.py
                         checks it against the real main call for call.

One controller iteration = two world steps of 32 ms (sup:705 and sup:754 -> 81-84).
"""
from __future__ import annotations

import math
import threading
import types
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import abi

BASIC_TIME_STEP_MS = 32          # Webots default; testing.wbt's WorldInfo is empty
WHEEL_RADIUS = 0.02              # E-puck.proto:100
AXLE_LENGTH = 0.052              # E-puck.proto:86
SEARCHING, HOMING = 1, 2         # RobotState, sup:180-183


# ------------------------------------------------------------------------------------------ devices
class _Motor:
    def __init__(self):
        self.velocity = 0.0

    def setPosition(self, _p):
        pass

    def setVelocity(self, v):
        self.velocity = float(v)

    def getVelocity(self):
        return self.velocity


class _DistanceSensor:
    def __init__(self, body, index):
        self._b, self._i = body, index

    def enable(self, _ms):
        pass

    def getValue(self):
        return self._b.ps[self._i]


class _Led:
    def set(self, _v):
        pass


class _Camera:
    """BGRA image; `red_pixels` of them saturated red while the robot is inside the food radius (the reference
    counts mask.sum() = 255 x pixels and grabs at 65000 < sum < 517140, sup:146-165, 609-620)."""
    W = H = 64

    def __init__(self, body, world):
        self._b, self._w = body, world

    def enable(self, _ms):
        pass

    def getWidth(self):
        return self.W

    def getHeight(self):
        return self.H

    def getImage(self):
        img = np.zeros((self.H, self.W, 4), np.uint8)
        if self._w.sees_food(self._b):
            img.reshape(-1, 4)[: self._w.red_pixels] = (0, 0, 255, 255)
        return img.tobytes()


class _Quiet:
    """emitter / receiver / display: accepted and ignored (rendering and radio are out of scope)"""

    def __getattr__(self, _name):
        return lambda *a, **k: None


class _Node:
    def __init__(self, getter):
        self._get = getter

    def getPosition(self):
        return list(self._get())

    def getField(self, _name):
        return self

    def getSFVec3f(self):
        return list(self._get())


class _Body:
    def __init__(self, name, pose, theta, ps):
        self.name = name
        self.pos = [float(v) for v in pose]
        self.theta = float(theta)
        self.ps = [float(ps)] * 8
        self.left, self.right = _Motor(), _Motor()
        self.stream = None       # recorded poses, one per controller iteration
        self.steps = 0           # world steps this robot's controller has taken


class _Supervisor:
    """what `Supervisor()` returns inside a controller process of robot `body`"""

    def __init__(self, world, body):
        self._w, self._b = world, body
        self._dev = {"left wheel motor": body.left, "right wheel motor": body.right,
                     "camera": _Camera(body, world)}
        for i in range(8):
            self._dev[f"ps{i}"] = _DistanceSensor(body, i)
        for i in range(10):
            self._dev[f"led{i}"] = _Led()

    def getBasicTimeStep(self):
        return float(BASIC_TIME_STEP_MS)

    def getDevice(self, name):
        return self._dev.setdefault(name, _Quiet())

    def getNumberOfDevices(self):
        return 0                  # no ground sensors are enumerated (sup:36-45 then leaves them None)

    def getName(self):
        return self._b.name

    def getTime(self):
        return self._w.time_of(self._b)

    def getSelf(self):
        return _Node(lambda: self._b.pos)

    def getFromDef(self, _defname):
        return _Node(lambda: self._w.food)

    def step(self, ms):
        return self._w.step(self._b, ms)

    def cleanup(self):
        pass


# ------------------------------------------------------------------------------------------ world
class HeadlessWorld:
    def __init__(self, food=(0.36, -0.15, 0.05), t_first: float = 0.0, food_radius: float = 0.08,
                 red_pixels: int = 400, max_iterations: Optional[int] = None, arena=None):
        """t_first = what getTime() returns after the FIRST step (the time the first controller iteration sees);
        after k steps the clock reads t_first + (k - 1) * 0.032 exactly (no accumulated rounding).
        arena = (xmin, xmax, ymin, ymax) walls for the integrated robots, or None."""
        self.food = [float(v) for v in food]
        self.t_first = float(t_first)
        self.food_radius = float(food_radius)
        self.red_pixels = int(red_pixels)
        self.max_iterations = max_iterations
        self.arena = arena
        self.bodies: Dict[str, _Body] = {}
        self.on_iteration: Optional[Callable[[str, int], None]] = None  # (robot, iteration) at the 2nd step
        self._barrier = None

    # -- robots
    def add_robot(self, name, pose, theta=0.0, ps=60.0, stream: Optional[Sequence] = None):
        b = _Body(name, pose, theta, ps)
        if stream is not None:
            b.stream = [list(map(float, p)) for p in stream]
        self.bodies[name] = b
        return b

    def supervisor(self, name) -> _Supervisor:
        return _Supervisor(self, self.bodies[name])

    def controller_module(self, name) -> types.ModuleType:
        """a module for sys.modules['controller'] whose Supervisor() is robot `name` of this world"""
        m = types.ModuleType("controller")
        world = self

        class Supervisor(_Supervisor):
            def __init__(self):
                super().__init__(world, world.bodies[name])

        class Node:
            DISTANCE_SENSOR = 40

        m.Supervisor = Supervisor
        m.Robot = Supervisor
        m.Node = Node
        for cls in ("Motor", "Field", "DistanceSensor", "LED", "Camera", "Emitter", "Receiver"):
            setattr(m, cls, type(cls, (), {}))
        return m

    def lockstep(self, n_threads: int):
        """several controller loops, one thread each: every world step waits for all of them (Webots does this
        between controller processes)"""
        self._barrier = threading.Barrier(n_threads)

    # -- clock and physics
    def time_of(self, b: _Body) -> float:
        return self.t_first + (b.steps - 1) * (BASIC_TIME_STEP_MS / 1000.0)

    def sees_food(self, b: _Body) -> bool:
        return math.hypot(b.pos[0] - self.food[0], b.pos[1] - self.food[1]) <= self.food_radius

    def step(self, b: _Body, ms) -> int:
        """one Webots step for robot b's controller: -1 when the run is over. Step 2i is the `while` step of
        controller iteration i (sup:705), step 2i + 1 the one at its end (sup:754)."""
        it, second = divmod(b.steps, 2)
        if second and self.on_iteration is not None:
            self.on_iteration(b.name, it)
        if self._barrier is not None:
            self._barrier.wait()
        if not second and self.max_iterations is not None and it >= self.max_iterations:
            return -1
        if b.stream is not None:
            if not second:                    # the pose this iteration will read
                if it >= len(b.stream):
                    return -1
                b.pos = list(b.stream[it])
        else:
            self._integrate(b, ms / 1000.0)
        b.steps += 1
        return 0

    def _integrate(self, b: _Body, dt: float):
        wl, wr = b.left.velocity, b.right.velocity
        v = WHEEL_RADIUS * (wl + wr) / 2.0            # Phase-1/drive_robot.py:36-47
        om = WHEEL_RADIUS * (wr - wl) / AXLE_LENGTH
        b.theta = math.atan2(math.sin(b.theta + om * dt), math.cos(b.theta + om * dt))
        x = b.pos[0] + v * dt * math.cos(b.theta)
        y = b.pos[1] + v * dt * math.sin(b.theta)
        if self.arena is not None:
            xmin, xmax, ymin, ymax = self.arena
            if x < xmin or x > xmax:
                x = min(max(x, xmin), xmax)
                b.theta = math.atan2(math.sin(math.pi - b.theta), math.cos(math.pi - b.theta))
            if y < ymin or y > ymax:
                y = min(max(y, ymin), ymax)
                b.theta = -b.theta
        b.pos[0], b.pos[1] = x, y


# ------------------------------------------------------------------------------------------ drop-in
class _NoLiveEntries:
    """pheromone_grids as Controller.startPheromone sees it (sup:594-607): the age-window evaporation already ran
    inside the engine's tick, so the Python loop over `.items()` finds nothing left to do. Works for any robot
    name, which also sidesteps the reference's leaked loop variable."""

    def __init__(self, algo):
        self._algo = algo

    def __getitem__(self, name):
        return {}

    def get(self, name, default=None):
        return {}

    def live(self, name) -> dict:
        """the real live entries {time: (intensity, [x, y, z])} read back from the engine"""
        return self._algo.engine.pheromone_grid(name)


class PheromoneAlgorithm:
    """Drop-in for the reference's PheromoneAlgorithm (sup:228-525) over a trail engine shared by all robots of a
    process. One engine tick (deposit -> sense -> adjust -> evaporate for every robot, sup:565-607) runs when the
    first robot of a world step calls deposit_pheromone; the other calls of that step return its results."""

    _shared: Dict[int, dict] = {}

    def __init__(self, robot, engine=None, world: Optional[HeadlessWorld] = None, robot_names=None,
                 start_delay: float = 5.0):
        self.robot = robot
        self.exploration_pheromone_intensity = 1.0      # sup:232-234
        self.higher_pheromone_intensity = 20.0
        self.homing_pheromone_intensity = 5.0
        self.evaporation_rate = 0.1                     # sup:241
        self.evaporation_interval = 2.0
        self.name = robot.supervisor.getName()
        self.start_position = self.get_robot_position()  # sup:248
        self.controller = None
        self.adjusted_left_speed = 0
        self.adjusted_right_speed = 0
        if engine is None:
            from .trail import TrailSwarm
            names = list(robot_names) if robot_names is not None else [self.name]
            if world is not None:
                starts = [world.bodies[n].pos for n in names]
                ps = world.bodies[names[0]].ps[0]
                food = world.food
            else:
                starts, ps, food = [self.start_position], robot.distance_sensors_values[0], self.get_food_position()
            engine = TrailSwarm(names, starts, food, ps=ps, start_delay=start_delay)
        self.engine = engine
        st = PheromoneAlgorithm._shared.setdefault(id(engine), {"lock": threading.Lock(), "t": None, "res": {},
                                                               "members": {}, "barrier": None})
        st["members"][self.name] = self
        self._st = st
        self.pheromone_grids = _NoLiveEntries(self)
        self.global_pheromone_grid = {}
        self._rec = None

    # -- the reference's small accessors
    def setController(self, controller):
        self.controller = controller

    def get_robot_position(self):                       # sup:255-258
        return self.robot.supervisor.getSelf().getPosition()

    def get_food_position(self):                        # sup:419-422
        return self.robot.supervisor.getFromDef("food").getField("translation").getSFVec3f()

    def lockstep(self, n_robots: int):
        """N controller threads share this engine: the tick runs once all N have reached deposit_pheromone (every
        robot's state and pose of this step are then final)"""
        self._st["barrier"] = threading.Barrier(n_robots)

    def _now(self):
        return self.controller.time["current_time"] if self.controller is not None else self.robot.supervisor.getTime()

    def _tick_all(self, t_now):
        st = self._st
        if st["barrier"] is not None:
            st["barrier"].wait()
        with st["lock"]:
            if st["t"] != t_now:
                pos = {}
                for name, m in st["members"].items():
                    homing = m.controller is not None and getattr(m.controller, "robot_state", SEARCHING) == HOMING
                    self.engine.set_state(name, homing)
                    pos[name] = m.get_robot_position()
                st["res"] = self.engine.startPheromone(list(pos), pos, t_now)
                st["t"] = t_now
        return st["res"][self.name]

    # -- the five methods
    def deposit_pheromone(self, robot_name, robot_pos):
        """sup:291-358; returns the (proxy of the) live grid like the reference does"""
        self._rec = self._tick_all(self._now())
        if self._rec["amount"] is not None:
            self.global_pheromone_grid.setdefault(robot_name, {})[self._rec["t"]] = (self._rec["amount"], robot_pos)
        return self.pheromone_grids

    def sense_pheromone(self, robot_name, robot_pos):
        """sup:379-417: {'current': [[t, (I, pos)], ...], 'front': ..., 'back': ..., 'left': ..., 'right': ...}"""
        s = (self._rec or {}).get("sense") or {k: [] for k in abi.DIR_NAMES}
        return {k: [[t, (inten, list(robot_pos))] for t, inten in v] for k, v in s.items()}

    def adjust_robot_behavior(self, sensed_pheromone, robot_name, robot_state=None):
        """sup:429-516 (3-argument form: experiment2.py:526)"""
        wl, wr = self._rec["speeds"]
        self.adjusted_left_speed, self.adjusted_right_speed = wl, wr
        return wl, wr

    def evaporate_pheromone(self, robot_name, elapsed_time, intensity, evaporation_rate=0.1):
        return intensity * math.exp(-evaporation_rate * elapsed_time)   # sup:519-525

    @property
    def last_record(self):
        return self._rec


# ------------------------------------------------------------------------------------------ main loop
def run_supervisor_loop(world: HeadlessWorld, engine, robot_names: Sequence[str], iterations: int,
                        start_delay: float = 5.0, on_tick: Optional[Callable[[int, float, dict], None]] = None):
    """Controller.main (sup:704-754) for every robot of `world` in one process, the pheromone path through ONE
    engine tick per iteration. Per robot and iteration, in the reference's order: reset speeds, read the proximity
    sensors, advance the clock (sup:711-718), Braitenberg base speeds (sup:721-725), SEARCHING -> HOMING when the
    camera shows food (sup:672-678, 609-620); SEARCHING: startPheromone from t >= 5 s, else the random walk
    (sup:731-744); HOMING: homing (sup:747-750); set the motors (sup:753); second world step (sup:754)."""
    names = list(robot_names)
    sup = {n: world.supervisor(n) for n in names}
    state = {n: SEARCHING for n in names}
    speeds = {n: [0.0, 0.0] for n in names}
    coeff = abi.BRAITENBERG_COEFFICIENTS
    log: List[dict] = []
    for it in range(iterations):
        if any(sup[n].step(BASIC_TIME_STEP_MS) == -1 for n in names):
            break
        t = sup[names[0]].getTime()
        active = []
        for n in names:
            b = world.bodies[n]
            sp = speeds[n]
            for i in range(2):
                sp[i] = 0.0
                for j in range(8):
                    sp[i] += coeff[j][i] * (1.0 - (b.ps[j] / (1024 / 2)))
            if state[n] == SEARCHING and world.sees_food(b):
                state[n] = HOMING
            if state[n] == HOMING or t >= start_delay:
                active.append(n)
            engine.set_state(n, state[n] == HOMING)
        res = {}
        if active:
            pos = {n: list(world.bodies[n].pos) for n in names}
            res = engine.startPheromone(names, pos, t)
        for n in names:
            b = world.bodies[n]
            if n in active and res[n]["speeds"] is not None:
                speeds[n][0], speeds[n][1] = res[n]["speeds"]
            b.left.setVelocity(speeds[n][0])
            b.right.setVelocity(speeds[n][1])
        if on_tick is not None:
            on_tick(it, t, res)
        log.append({"t": t, "records": res, "speeds": {n: list(speeds[n]) for n in names}})
        if any(sup[n].step(BASIC_TIME_STEP_MS) == -1 for n in names):
            break
    return log
