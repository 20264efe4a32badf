"""Row-strip sharding of the arena across GPUs (no reference counterpart).

Rank g owns global rows [row0, row1) of every channel plus one ghost row on each side, and the
agents whose current cell lies in those rows. One tick:

    1. deposit            (local: agents only deposit into rows they own)
    2. halo exchange      own first/last interior row -> neighbour's ghost row, per channel
                          (the interior rows' stencil runs while the halo is in flight)
    3. evaporate+diffuse  edge rows after the halo landed; sense/decide/move
    4. migration          agents whose row left the strip are packed (ordered by slot), sent to
                          the neighbour, and placed into its empty slots (lowest first)

The field result is bit-identical to the unsharded run; agent records are too (each agent's update
depends only on its own record and on field values). The host logic is written against a small
backend interface so it can be exercised on CPU (gloo, world_size 2) with the test oracle as the
compute backend; the product backend is CudaStrip (libpherofield.so).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import abi


def strip_bounds(H: int, world: int, rank: int):
    """contiguous, balanced row strips: the first H % world strips get one extra row"""
    base, extra = divmod(H, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def strip_config(cfg: abi.PFConfig, world: int, rank: int, device: Optional[int] = None) -> abi.PFConfig:
    c = cfg.copy()
    c.row0, c.row1 = strip_bounds(cfg.H, world, rank)
    if device is not None:
        c.device = device
    return c


def owner_of_rows(H: int, world: int, rows: np.ndarray) -> np.ndarray:
    bounds = np.array([strip_bounds(H, world, r)[1] for r in range(world)])
    return np.searchsorted(bounds, rows, side="right")


def agent_rows(cfg: abi.PFConfig, x: np.ndarray) -> np.ndarray:
    """global row of each agent, with the library's truncation rule in float32"""
    inv = np.float32(1.0) / np.float32(cfg.cell_size)
    fr = (np.asarray(x, np.float32) - np.float32(cfg.origin_x)) * inv
    return np.clip(np.trunc(fr), 0, cfg.H - 1).astype(np.int64)


class CudaStrip:
    """Product backend: one PheroField strip + its Swarm on one GPU."""

    def __init__(self, cfg: abi.PFConfig, x, y, theta, state, gid, capacity: int, migrate_cap: int):
        from .field import PheroField, Swarm
        self.cfg = cfg
        self.field = PheroField(cfg)
        self.swarm = Swarm(x, y, theta, state, device=self.field.device, capacity=capacity, gid=gid)
        if len(np.atleast_1d(x)) < capacity:  # unused slots are holes
            self.swarm.meta[len(np.atleast_1d(x)):] = abi.STATE_EMPTY
        self.rows = cfg.row1 - cfg.row0
        self.mcap = migrate_cap
        words = abi.PF_MIGRATE_WORDS * (migrate_cap + 1)
        dev = self.field.device
        self.send_up = torch.zeros(words, dtype=torch.int32, device=dev)
        self.send_down = torch.zeros(words, dtype=torch.int32, device=dev)
        self.recv_up = torch.zeros(words, dtype=torch.int32, device=dev)    # from rank-1
        self.recv_down = torch.zeros(words, dtype=torch.int32, device=dev)  # from rank+1
        self.field.reserve_agents(capacity)

    # halo tensors: lists of contiguous [W] rows, one per channel. The two buffers never move, so
    # the row views of both are built once and selected by the swap parity.
    def _rows_cache(self):
        if getattr(self, "_cache", None) is None:
            C, rows = self.cfg.C, self.rows
            both = []
            for which in (0, 1):
                v = self.field.view(which, with_ghosts=True)
                both.append({"edge": ([v[c, 1] for c in range(C)], [v[c, rows] for c in range(C)]),
                             "ghost": ([v[c, 0] for c in range(C)], [v[c, rows + 1] for c in range(C)])})
            self._cache = both
            self._parity = 0
        return self._cache[self._parity]

    def edge_rows(self):
        return self._rows_cache()["edge"]

    def ghost_rows(self):
        return self._rows_cache()["ghost"]

    def deposit(self):
        self.field.agents_deposit(self.swarm)

    def field_rows(self, r0, r1):
        self.field.step_field_rows(r0, r1)

    def agents(self, flags):
        self.field.agents_step(self.swarm, flags)

    def swap(self):
        self.field.swap()
        if getattr(self, "_cache", None) is not None:
            self._parity ^= 1

    def sort_agents(self):
        self.field.sort_agents(self.swarm)

    def pack(self):
        self.field.migrate_pack(self.swarm, self.send_up, self.send_down, self.mcap)

    def unpack(self, from_up: bool, from_down: bool):
        if from_up:
            self.field.migrate_unpack(self.swarm, self.recv_up, self.mcap)
        if from_down:
            self.field.migrate_unpack(self.swarm, self.recv_down, self.mcap)

    def synchronize(self):
        torch.cuda.synchronize(self.field.device)


class StripWorker:
    """Phase logic of one rank, independent of how bytes move between ranks."""

    def __init__(self, backend, rank: int, world: int):
        self.b = backend
        self.rank, self.world = rank, world
        self.has_up = rank > 0
        self.has_down = rank < world - 1
        self.ticks = 0

    def flags(self, sense_enabled: bool, move: bool = True) -> int:
        return (abi.STEP_MOVE if move else 0) | (abi.STEP_SENSE if sense_enabled else 0)

    # phase 1
    def deposit(self, sense_enabled: bool):
        if sense_enabled:
            self.b.deposit()

    # phase 2: (peer, tensors) pairs
    def halo_sends(self):
        top, bottom = self.b.edge_rows()
        out = []
        if self.has_up:
            out.append((self.rank - 1, top))
        if self.has_down:
            out.append((self.rank + 1, bottom))
        return out

    def halo_recvs(self):
        top, bottom = self.b.ghost_rows()
        out = []
        if self.has_up:
            out.append((self.rank - 1, top))
        if self.has_down:
            out.append((self.rank + 1, bottom))
        return out

    def field_interior(self):
        if self.b.rows > 2:
            self.b.field_rows(1, self.b.rows - 1)

    def field_edges_and_agents(self, sense_enabled: bool, move: bool = True):
        rows = self.b.rows
        self.b.field_rows(0, min(1, rows))
        if rows > 1:
            self.b.field_rows(rows - 1, rows)
        self.b.agents(self.flags(sense_enabled, move))
        self.b.swap()

    # phase 4
    def pack(self):
        self.b.pack()

    def migrate_sends(self):
        out = []
        if self.has_up:
            out.append((self.rank - 1, [self.b.send_up]))
        if self.has_down:
            out.append((self.rank + 1, [self.b.send_down]))
        return out

    def migrate_recvs(self):
        out = []
        if self.has_up:
            out.append((self.rank - 1, [self.b.recv_up]))
        if self.has_down:
            out.append((self.rank + 1, [self.b.recv_down]))
        return out

    def unpack(self):
        self.b.unpack(self.has_up, self.has_down)
        self.ticks += 1


class LocalDriver:
    """All strips in one process (one or several devices): bytes move with Tensor.copy_.
    Used for k-strips-equals-one-GPU parity tests and as the reference for DistDriver."""

    def __init__(self, workers: Sequence[StripWorker]):
        self.w = list(workers)

    @staticmethod
    def _move(workers, sends_of, recvs_of):
        inbox = {}
        for w in workers:
            for peer, tensors in sends_of(w):
                inbox[(w.rank, peer)] = tensors
        for w in workers:
            for peer, tensors in recvs_of(w):
                src = inbox[(peer, w.rank)]
                for d, s in zip(tensors, src):
                    d.copy_(s)

    def tick(self, sense_enabled: bool = True):
        for w in self.w:
            w.deposit(sense_enabled)
        self._move(self.w, lambda w: w.halo_sends(), lambda w: w.halo_recvs())
        for w in self.w:
            w.field_interior()
            w.field_edges_and_agents(sense_enabled)
        for w in self.w:
            w.pack()
        self._move(self.w, lambda w: w.migrate_sends(), lambda w: w.migrate_recvs())
        for w in self.w:
            w.unpack()


class DistDriver:
    """One process per GPU over torch.distributed (NCCL on GPUs, gloo on CPU for tests).

    The halo travels on a side stream while the interior rows' stencil runs; only the two edge
    rows and the agent step wait for it."""

    def __init__(self, worker: StripWorker, group=None, overlap: bool = True):
        import torch.distributed as dist
        self.dist = dist
        self.w = worker
        self.group = group
        self.cuda = torch.cuda.is_available() and dist.get_backend(group) == "nccl"
        self.overlap = overlap and self.cuda
        if self.cuda:
            self.comm_stream = torch.cuda.Stream()
            self.ev_ready = torch.cuda.Event()
            self.ev_done = torch.cuda.Event()

    def _exchange(self, sends, recvs):
        dist = self.dist
        ops = []
        for peer, tensors in recvs:
            for t in tensors:
                ops.append(dist.P2POp(dist.irecv, t, peer, self.group))
        for peer, tensors in sends:
            for t in tensors:
                ops.append(dist.P2POp(dist.isend, t, peer, self.group))
        if not ops:
            return []
        return dist.batch_isend_irecv(ops)

    def tick(self, sense_enabled: bool = True, move: bool = True):
        """move=False: the caller owns the poses (pf_agents_observe); nothing migrates"""
        w = self.w
        w.deposit(sense_enabled)
        if self.overlap:
            main = torch.cuda.current_stream()
            self.ev_ready.record(main)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(self.ev_ready)
                reqs = self._exchange(w.halo_sends(), w.halo_recvs())
                for r in reqs:
                    r.wait()
                self.ev_done.record(self.comm_stream)
            w.field_interior()
            main.wait_event(self.ev_done)
        else:
            for r in self._exchange(w.halo_sends(), w.halo_recvs()):
                r.wait()
            w.field_interior()
        w.field_edges_and_agents(sense_enabled, move)
        if not move:
            w.ticks += 1
            return
        w.pack()
        for r in self._exchange(w.migrate_sends(), w.migrate_recvs()):
            r.wait()
        w.unpack()


def partition_agents(cfg: abi.PFConfig, world: int, x, y, theta, state=None, gid=None):
    """split initial agents by the owner strip of their row; returns per-rank dicts"""
    x = np.asarray(x, np.float32)
    n = len(x)
    state = np.full(n, abi.STATE_SEARCHING, np.int64) if state is None else np.broadcast_to(
        np.asarray(state, np.int64), (n,))
    gid = np.arange(n, dtype=np.uint32) if gid is None else np.asarray(gid, np.uint32)
    owner = owner_of_rows(cfg.H, world, agent_rows(cfg, x))
    parts = []
    for r in range(world):
        idx = np.nonzero(owner == r)[0]
        parts.append({"x": x[idx], "y": np.asarray(y, np.float32)[idx],
                      "theta": np.asarray(theta, np.float32)[idx], "state": state[idx],
                      "gid": gid[idx]})
    return parts
