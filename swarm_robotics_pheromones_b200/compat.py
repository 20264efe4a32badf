"""ReferenceCompat — the reference's supervisor-level calls over N robots on the B200 field.

Mirrors the method names, argument orders and return shapes of the reference's
PheromoneAlgorithm / Controller (Phase-3/pheromone_algo_supervisor.py, `sup` below) so that
Controller.main-style code can call it unchanged:

    deposit_pheromone(robot_name, robot_pos)                     sup:291
    sense_pheromone(robot_name, robot_pos) -> {'current': ...}   sup:379
    adjust_robot_behavior(sensed_pheromone, robot_name)          sup:429   (3-arg: experiment2.py:526)
    evaporate_pheromone(robot_name, elapsed_time, intensity, evaporation_rate=0.1)   sup:519
    startPheromone(robot_names) / homing(robot_names)            sup:565 / sup:627

plus batched forms (`tick_host`) that take every robot's pose at once from pinned host memory —
the call a Webots supervisor driving many robots makes once per tick.

Differences that follow from the dense field: the sensed lists hold the FIELD value of
the cell in each direction (one item per direction that lies inside the arena) instead of
timestamped trail entries, and evaporation is the field's per-tick keep factor. Errors: unknown
robot names raise KeyError like the reference's dict lookups do.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence

import numpy as np
import torch

from . import abi
from .field import PheroField, Swarm


class ReferenceCompat:
    def __init__(self, cfg: abi.PFConfig, robot_names: Sequence[str], start_positions,
                 headings=None, states=None, start_delay: float = 5.0, time0: float = 0.0):
        field = PheroField(cfg)
        names = list(robot_names)
        pos = np.asarray(start_positions, dtype=np.float32).reshape(len(names), -1)
        n = len(names)
        th = np.zeros(n, np.float32) if headings is None else np.asarray(headings, np.float32)
        st = abi.STATE_SEARCHING if states is None else states
        swarm = Swarm(pos[:, 0], pos[:, 1], th, st, device=field.device)
        # first observation sets has_moved = True (old_pos is None, sup:261-262)
        field.set_time(time0, start_delay)
        self._attach(field, swarm, names)

    @classmethod
    def from_field(cls, field: PheroField, swarm: Swarm, robot_names: Optional[Sequence[str]] = None):
        """wrap an existing field + swarm (robot i is called str(i) unless names are given)"""
        self = object.__new__(cls)
        self._attach(field, swarm, robot_names)
        return self

    def _attach(self, field: PheroField, swarm: Swarm, names):
        self.field = field
        self.cfg = field.cfg
        self.swarm = swarm
        n = swarm.n
        self.names = list(names) if names is not None else None
        self.index = _NameIndex(self.names, n)
        self.field.reserve_agents(n)
        dev = self.field.device
        # pinned staging for the batched host path
        self._h_pose = torch.empty((3, n), dtype=torch.float32).pin_memory()
        self._d_pose = torch.empty((3, n), dtype=torch.float32, device=dev)
        self._h_out = torch.empty((2, n), dtype=torch.float32).pin_memory()
        self._d_dec = torch.empty(n, dtype=torch.uint8, device=dev)
        self._h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
        self.h2d_bytes = self._h_pose.numel() * 4
        self.d2h_bytes = self._h_out.numel() * 4 + self._h_dec.numel()
        self._last_sense = None

    # ---- helpers ---------------------------------------------------------------------------
    def _sel(self, robot_name) -> Optional[int]:
        if robot_name is None:
            return None
        return self.index[robot_name]  # KeyError for unknown names, like the reference's dicts

    def set_state(self, robot_name, homing: bool):
        i = self.index[robot_name]
        st = abi.STATE_HOMING if homing else abi.STATE_SEARCHING
        m = int(self.swarm.meta[i].item()) & 0xFFFFFFFF
        self.swarm.meta[i] = int(np.array((m & ~abi.META_STATE_MASK) | st, np.uint32).view(np.int32))

    def _observe_one(self, i: int, robot_pos):
        n = self.swarm.n
        x = self.swarm.x.clone()
        y = self.swarm.y.clone()
        x[i] = float(robot_pos[0])
        y[i] = float(robot_pos[1])
        # only robot i gets a new observation; the others keep pose and moved flag
        meta_before = self.swarm.meta.clone()
        self.field.observe(self.swarm, x, y)
        keep = torch.ones(n, dtype=torch.bool, device=self.swarm.device)
        keep[i] = False
        self.swarm.meta[keep] = meta_before[keep]

    # ---- the reference's five methods --------------------------------------------------------
    def deposit_pheromone(self, robot_name, robot_pos):
        """sup:291-358 for one robot: store the pose, gate on has_moved, deposit by state."""
        i = self._sel(robot_name)
        self._observe_one(i, robot_pos)
        # deposit only this robot: mask the others' moved flag for the call
        meta = self.swarm.meta.clone()
        others = torch.ones(self.swarm.n, dtype=torch.bool, device=self.swarm.device)
        others[i] = False
        self.swarm.meta[others] = meta[others] & ~abi.META_MOVED
        self.field.agents_deposit(self.swarm)
        self.swarm.meta[others] = meta[others]
        return self.field  # the reference returns its live grid object

    def sense_pheromone(self, robot_name, robot_pos):
        """sup:379-417: dict in the reference's key order; each present direction holds one item
        [time, (intensity, [x, y, z])]."""
        i = self._sel(robot_name)
        x = torch.tensor([float(robot_pos[0])], device=self.swarm.device)
        y = torch.tensor([float(robot_pos[1])], device=self.swarm.device)
        th = self.swarm.theta[i:i + 1]
        v = self.field.sense(x, y, th, self.swarm.meta[i:i + 1]).cpu().numpy()[0]
        t, _ = self.field.time()
        out = {k: [] for k in abi.DIR_NAMES}
        for d, k in enumerate(abi.DIR_NAMES):
            if not np.isnan(v[d]):
                out[k].append([t, (float(v[d]), list(robot_pos))])
        self._last_sense = (i, v)
        return out

    def adjust_robot_behavior(self, sensed_pheromone, robot_name, robot_state=None):
        """sup:429-516: wheel speeds for one robot from the device decide kernel (the sensed dict is
        re-gathered on the device from the same field, so it need not be passed back down)."""
        i = self._sel(robot_name)
        if robot_state is not None:
            self.set_state(robot_name, robot_state == abi.STATE_HOMING)
        # decide only (no move): run the kernel on a 1-agent view of the SoA
        view = _SwarmSlice(self.swarm, i)
        self.field.agents_step(view, abi.STEP_SENSE)
        return float(self.swarm.wl[i].item()), float(self.swarm.wr[i].item())

    def evaporate_pheromone(self, robot_name, elapsed_time, intensity, evaporation_rate=0.1):
        """sup:519-525: intensity * exp(-rate * elapsed); scalars or arrays."""
        scalar = np.isscalar(elapsed_time) and np.isscalar(intensity)
        dev = self.field.device
        e = torch.as_tensor(np.atleast_1d(np.asarray(elapsed_time, np.float64)), device=dev)
        it = torch.as_tensor(np.atleast_1d(np.asarray(intensity, np.float64)), device=dev)
        e, it = torch.broadcast_tensors(e, it)
        out = self.field.evaporate_trail(e.contiguous(), it.contiguous(), float(evaporation_rate))
        out = out.cpu().numpy()
        return float(out[0]) if scalar else out

    # ---- tick drivers -------------------------------------------------------------------------
    def startPheromone(self, robot_names: Optional[Iterable[str]] = None, positions=None):
        """sup:565-607 for the named robots (all when None): deposit -> sense -> adjust ->
        evaporate. `positions` = {name: pos} or an [N,2+] array for all robots."""
        return self._tick(robot_names, positions)

    def homing(self, robot_names: Optional[Iterable[str]] = None, positions=None):
        """sup:627-663: same tick; robots must be in HOMING state (set_state)."""
        return self._tick(robot_names, positions)

    def _tick(self, robot_names, positions):
        n = self.swarm.n
        pose = self._h_pose
        cur = self.swarm.to_numpy()
        pose[0].copy_(torch.from_numpy(cur["x"]))
        pose[1].copy_(torch.from_numpy(cur["y"]))
        pose[2].copy_(torch.from_numpy(cur["theta"]))
        if positions is not None:
            if isinstance(positions, dict):
                for name, p in positions.items():
                    i = self.index[name]
                    pose[0, i], pose[1, i] = float(p[0]), float(p[1])
            else:
                p = np.asarray(positions, np.float32).reshape(n, -1)
                pose[0].copy_(torch.from_numpy(np.ascontiguousarray(p[:, 0])))
                pose[1].copy_(torch.from_numpy(np.ascontiguousarray(p[:, 1])))
        if robot_names is not None:
            for name in robot_names:
                self.index[name]
        wl, wr, dec = self.tick_host(pose)
        if robot_names is None:
            names = self.names if self.names is not None else [str(i) for i in range(n)]
        else:
            names = list(robot_names)
        return {name: (float(wl[self.index[name]]), float(wr[self.index[name]])) for name in names}

    def tick_host(self, pose_host: torch.Tensor):
        """One supervisor tick for every robot from HOST poses (pinned [3][N]: x, y, theta):
        H2D poses -> observe -> deposit -> sense/decide -> evaporate(+diffuse) -> D2H wheel speeds
        and decisions. The caller (Webots in the reference) owns the motion, so PF_STEP_MOVE is
        not run. Returns (wl, wr, decision) host tensors (views of pinned staging) as soon as they have
        arrived: the supervisor needs the speeds, not the field — the evaporate+diffuse kernel of this
        tick may still be running and everything issued later is ordered behind it on the stream."""
        self.tick_host_async(pose_host)
        self._ev_out.synchronize()
        return self._h_out[0], self._h_out[1], self._h_dec

    CHUNKS = 4

    def _chunk_views(self):
        """robot sub-ranges as pf_agents views (pointer arithmetic only), for the pipelined host path"""
        if getattr(self, "_chunks", None) is None:
            n = self.swarm.n
            k = self.CHUNKS if n >= (1 << 16) else 1
            cuts = [(n * i // k) // 256 * 256 for i in range(k)] + [n]
            self._chunks = [(a, b, _SwarmSlice(self.swarm, a, b - a)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        return self._chunks

    def tick_host_async(self, pose_host: torch.Tensor):
        """The tick as a pipeline over robot sub-ranges (the copies dominate: 12 B in, 9 B out per robot):
          copy-in stream   x and y of every chunk first, the headings behind them
          main stream      observe chunk i as soon as its x / y have landed; deposit (all robots: it needs no
                           heading) beside the headings' H2D; then decide chunk i -> evaporate+diffuse
          copy-out stream  D2H of chunk i's speeds and decisions beside decide chunk i+1, the rest of the headings'
                           H2D (PCIe is full duplex) and the field kernel
        Before the start delay observe is skipped: the reference only stores poses inside deposit_pheromone
        (sup:279-289, 577), so the first deposit after t = 5 s is unconditional (old_pos is None, sup:261-262)."""
        f, sw = self.field, self.swarm
        main = torch.cuda.current_stream(f.device)
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream(f.device)       # copy-out
            self._cin = torch.cuda.Stream(f.device)        # copy-in
            self._ev_in = [torch.cuda.Event() for _ in range(self.CHUNKS)]
            self._ev_th = [torch.cuda.Event() for _ in range(self.CHUNKS)]
            self._ev_dec = [torch.cuda.Event() for _ in range(self.CHUNKS)]
            self._ev_out = torch.cuda.Event()
            self._ev_free = torch.cuda.Event()
            self._ev_free.record(main)
        t, _ = f.time()
        sensing = t >= f.start_delay
        chunks = self._chunk_views()
        self._cin.wait_event(self._ev_free)                # the previous tick has consumed _d_pose
        with torch.cuda.stream(self._cin):
            for i, (a, b, _) in enumerate(chunks):
                for r in range(2):  # row by row: contiguous 1-D copies (a strided 2-D one is staged by torch)
                    self._d_pose[r, a:b].copy_(pose_host[r, a:b], non_blocking=True)
                self._ev_in[i].record(self._cin)
            for i, (a, b, _) in enumerate(chunks):
                self._d_pose[2, a:b].copy_(pose_host[2, a:b], non_blocking=True)
                self._ev_th[i].record(self._cin)
        for i, (a, b, view) in enumerate(chunks):
            main.wait_event(self._ev_in[i])
            if sensing:                                    # storingPosition + has_moved; the heading follows below
                f.observe(view, self._d_pose[0, a:], self._d_pose[1, a:], None)
            else:                                          # poses only; no has_moved bookkeeping yet
                sw.x[a:b].copy_(self._d_pose[0, a:b]); sw.y[a:b].copy_(self._d_pose[1, a:b])
        if sensing:
            f.agents_deposit(sw)
        self._side.wait_event(self._ev_out)                # (re-use of the staging: previous D2H done)
        for i, (a, b, view) in enumerate(chunks):
            main.wait_event(self._ev_th[i])
            sw.theta[a:b].copy_(self._d_pose[2, a:b])
            # decide on the main stream (two HBM-bound kernels side by side only slow each other down);
            # the kernel writes the u8 decisions itself
            f.agents_step(view, abi.STEP_SENSE if sensing else 0, decisions=self._d_dec[a:])
            self._ev_dec[i].record(main)
            with torch.cuda.stream(self._side):            # D2H rides beside the next chunk and the stencil
                self._side.wait_event(self._ev_dec[i])
                self._h_out[0, a:b].copy_(sw.wl[a:b], non_blocking=True)
                self._h_out[1, a:b].copy_(sw.wr[a:b], non_blocking=True)
                self._h_dec[a:b].copy_(self._d_dec[a:b], non_blocking=True)
        self._ev_free.record(main)
        self._ev_out.record(self._side)
        f.step(sw, 1, abi.STEP_FIELD)  # evaporate+diffuse only; advances the tick clock
        main.wait_event(self._ev_out)

    def close(self):
        self.field.close()


class _NameIndex:
    """name -> slot; unnamed swarms use str(i) / int i. Unknown names raise KeyError like the
    reference's per-robot dicts."""

    def __init__(self, names, n):
        self.n = n
        self.map = None if names is None else {nm: i for i, nm in enumerate(names)}

    def __getitem__(self, name):
        if self.map is not None:
            return self.map[name]
        i = int(name)
        if not 0 <= i < self.n:
            raise KeyError(name)
        return i


class _SwarmSlice:
    """pf_agents view of `n` slots of a Swarm starting at slot i (pointer arithmetic only)."""

    def __init__(self, sw: Swarm, i: int, n: int = 1):
        s = abi.PFAgents()
        s.n = n
        s.x, s.y, s.theta = sw.x[i:].data_ptr(), sw.y[i:].data_ptr(), sw.theta[i:].data_ptr()
        s.wl, s.wr = sw.wl[i:].data_ptr(), sw.wr[i:].data_ptr()
        s.meta, s.gid = sw.meta[i:].data_ptr(), sw.gid[i:].data_ptr()
        self._s = s

    def struct(self):
        return self._s
