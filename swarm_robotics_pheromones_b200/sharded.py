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

from typing import Optional, Sequence

import numpy as np
import torch

from . import abi


def strip_bounds(H: int, world: int, rank: int):
    """contiguous, balanced row strips: the first H % world strips get one extra row"""
    base, extra = divmod(H, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def weighted_bounds(row_cost: np.ndarray, world: int, min_rows: int = 128):
    """strip boundaries [b0 = 0, b1, ..., b_world = H] that give every strip about the same total of `row_cost`
    (bytes a tick moves for that row: its cells plus the robots standing on it) — for populations that crowd a
    few rows, like config 4's nest. Every strip keeps at least `min_rows` rows."""
    cost = np.asarray(row_cost, np.float64)
    H = len(cost)
    assert world * min_rows <= H
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    b = [0]
    for r in range(1, world):
        want = cum[-1] * r / world
        k = int(np.searchsorted(cum, want))
        k = max(k, b[-1] + min_rows)
        k = min(k, H - (world - r) * min_rows)
        b.append(k)
    b.append(H)
    return b


def strip_config(cfg: abi.PFConfig, world: int, rank: int, device: Optional[int] = None,
                 bounds: Optional[Sequence[int]] = None) -> abi.PFConfig:
    c = cfg.copy()
    c.row0, c.row1 = (bounds[rank], bounds[rank + 1]) if bounds is not None else strip_bounds(cfg.H, world, rank)
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
        cfg = cfg.copy()
        cfg.migrate_cap = migrate_cap  # the peer mailbox the neighbours pack their leavers into (pf_strip_tick)
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

    def tile_rows(self):
        """local rows whose next-tick deposits the stencil can apply (pf_tile_rows); (0, 0) = none"""
        return self.field.tile_rows(self.swarm)

    def swap(self):
        self.field.swap()
        if getattr(self, "_cache", None) is not None:
            self._parity ^= 1

    def sort_agents(self):
        self.field.sort_agents(self.swarm)

    def pack(self):
        self.field.migrate_pack(self.swarm, self.send_up, self.send_down, self.mcap)

    def unpack(self, from_up: bool, from_down: bool):
        # one pass: arrivals from the upper neighbour first, then from the lower one
        if from_up or from_down:
            self.field.migrate_unpack(self.swarm, self.recv_up if from_up else None, self.mcap,
                                      self.recv_down if from_down else None)

    def synchronize(self):
        torch.cuda.synchronize(self.field.device)

    # ---- peer-store halo (NVLink P2P): the halo exchange fused into the deposit / stencil kernels
    peer = False

    def peer_attach_local(self, up: Optional["CudaStrip"], down: Optional["CudaStrip"]):
        """neighbour strips living in this process on the same device: raw pointers"""
        self.field.peer_attach(up.field.peer_export() if up is not None else None,
                               down.field.peer_export() if down is not None else None)
        self.peer = True

    def peer_attach_ipc(self, rank: int, world: int, group=None):
        """one process per GPU: exchange CUDA IPC handles of the field buffers and the flag words with
        the two neighbour ranks and map them (cudaIpcOpenMemHandle enables NVLink peer access)"""
        import torch.distributed as dist
        info = self.field.peer_export()
        mine = (self.field.peer_ipc_handles(), int(info.rows), int(info.pitch), int(info.chan_stride),
                int(info.mig_cap))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)

        def opened(r):
            if r < 0 or r >= world:
                return None
            handles, rows, pitch, cs, mig_cap = everyone[r]
            dims = abi.PFPeerInfo()
            dims.rows, dims.pitch, dims.chan_stride, dims.mig_cap = rows, pitch, cs, mig_cap
            return self.field.peer_ipc_open(handles, dims)

        self.field.peer_attach(opened(rank - 1), opened(rank + 1))
        self.peer = True

    def peer_prime(self):
        """make the neighbours' ghost rows of the current buffer valid (initial state / after uploads)"""
        self.field.peer_push_edges()
        self.field.peer_signal()

    def peer_signal(self):
        self.field.peer_signal()

    def peer_wait(self):
        self.field.peer_wait()

    def strip_tick(self, flags: int, phases: int = abi.TICK_ALL):
        """the whole tick by one C call (pf_strip_tick); needs the peers attached (their mailboxes take the
        leavers). Used by DistDriver / LocalDriver whenever `peer` is set."""
        self.field.strip_tick(self.swarm, flags, phases)


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

    # Deposits riding on the stencil (what pf_step does by itself on one GPU): the move kernel hands
    # the next deposit of the robots in the interior rows `dep` to the stencil launch over exactly
    # those rows. hand_over=False on the tick before a re-sort, or if the next tick does not deposit.
    dep = (0, 0)

    def begin_tick(self, hand_over: bool = True):
        self.dep = (0, 0)
        if hand_over and hasattr(self.b, "tile_rows"):
            self.dep = tuple(self.b.tile_rows())

    # phase 3: the stencil is split so that both exchanges hide behind it —
    #   interior part A  (while the halo is in flight)
    #   agents           (needs the ghost rows; produces the leavers)
    #   interior part B + the two edge rows (while the migration buffers are in flight)
    def _split(self):
        rows = self.b.rows
        lo, hi = 1, max(1, rows - 1)
        if self.dep[1] > self.dep[0]:
            return lo, self.dep[0], hi
        return lo, lo + (hi - lo) // 2, hi

    def field_interior_a(self):
        lo, mid, _ = self._split()
        if mid > lo:
            self.b.field_rows(lo, mid)

    def agents(self, sense_enabled: bool, move: bool = True):
        flags = self.flags(sense_enabled, move)
        if self.dep[1] > self.dep[0] and sense_enabled and move:
            flags |= abi.STEP_HANDOVER
        else:
            self.dep = (0, 0)
        self.b.agents(flags)

    def field_rest(self):
        rows = self.b.rows
        _, mid, hi = self._split()
        last_done = False
        if self.dep[1] > self.dep[0]:
            d0, d1 = self.dep
            self.b.field_rows(d0, d1)          # ONE call over exactly these rows: the deposits ride along
            last_done = True
            if d1 < rows:                       # the band below it, last row included (one launch)
                self.b.field_rows(d1, rows)
        elif hi > mid:
            self.b.field_rows(mid, hi)
        self.b.field_rows(0, min(1, rows))
        if rows > 1 and not last_done:
            self.b.field_rows(rows - 1, rows)
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

    def attach_peers(self):
        """fuse the halo into the kernels: every strip stores its edge rows straight into its
        neighbours' ghost rows (all strips of this driver live on one device)"""
        for i, w in enumerate(self.w):
            w.b.peer_attach_local(self.w[i - 1].b if i > 0 else None,
                                  self.w[i + 1].b if i + 1 < len(self.w) else None)
        for w in self.w:
            w.b.peer_prime()
        for w in self.w:
            w.b.peer_wait()
        self.peer = True

    peer = False

    def tick(self, sense_enabled: bool = True, hand_over: bool = False):
        """hand_over=True lets the interior rows' NEXT deposits ride on this tick's stencil: only when another
        depositing tick follows before the field, the counters or the agents are read or re-sorted (run() decides
        this by itself)."""
        if self.peer and all(hasattr(w.b, "strip_tick") for w in self.w):
            flags = strip_flags(sense_enabled, True, hand_over)
            for phase in (abi.TICK_A, abi.TICK_B, abi.TICK_C):  # one stream: every strip's signal before any wait
                for w in self.w:
                    w.b.strip_tick(flags, phase)
            for w in self.w:
                w.ticks += 1
            return
        for w in self.w:
            w.begin_tick(hand_over and sense_enabled)
            w.deposit(sense_enabled)
            if self.peer:
                w.b.peer_signal()
        if not self.peer:
            self._move(self.w, lambda w: w.halo_sends(), lambda w: w.halo_recvs())
        for w in self.w:
            w.field_interior_a()
            if self.peer:
                w.b.peer_wait()
            w.agents(sense_enabled)
            w.pack()
        self._move(self.w, lambda w: w.migrate_sends(), lambda w: w.migrate_recvs())
        for w in self.w:
            w.field_rest()
            w.unpack()


class DistDriver:
    """One process per GPU over torch.distributed (NCCL on GPUs, gloo on CPU for tests).

    The halo travels on a side stream while the interior rows' stencil runs; only the two edge
    rows and the agent step wait for it."""

    def __init__(self, worker: StripWorker, group=None, overlap: bool = True, profile: bool = False,
                 legacy: bool = False):
        """legacy=True keeps the phase-by-phase host loop with NCCL migration (the round-1 path; also what
        profile=True uses, for its per-phase events)"""
        import torch.distributed as dist
        self.profile = profile
        self.legacy = legacy
        self._marks = []  # per tick: list of (label, event) on the main stream
        self.dist = dist
        self.w = worker
        self.group = group
        self.cuda = torch.cuda.is_available() and dist.get_backend(group) == "nccl"
        self.overlap = overlap and self.cuda
        if self.cuda:
            self.comm_stream = torch.cuda.Stream()
            self.ev_ready = torch.cuda.Event()
            self.ev_done = torch.cuda.Event()
            self.ev_pack = torch.cuda.Event()
            self.ev_mig = torch.cuda.Event()

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

    def _comm(self, sends, recvs, ev_before, ev_after):
        """run one neighbour exchange on the comm stream, fenced by events against the main stream"""
        main = torch.cuda.current_stream()
        ev_before.record(main)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev_before)
            for r in self._exchange(sends, recvs):
                r.wait()
            ev_after.record(self.comm_stream)

    def tick(self, sense_enabled: bool = True, move: bool = True, hand_over: bool = False):
        """move=False: the caller owns the poses (pf_agents_observe); nothing migrates.
        hand_over=True lets the interior rows' NEXT deposits ride on this tick's stencil: only when another
        depositing tick follows before the field, the counters or the agents are read or re-sorted (run() decides
        this by itself)."""
        w = self.w
        if getattr(w.b, "peer", False) and hasattr(w.b, "strip_tick") and not self.profile and not self.legacy:
            # one C call enqueues the whole tick; halo rows and leavers cross as peer stores + epoch flags
            w.b.strip_tick(strip_flags(sense_enabled, move, hand_over))
            w.ticks += 1
            return
        w.begin_tick(hand_over and sense_enabled and move)
        marks = [] if (self.profile and self.cuda) else None

        def mark(label):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream())
                marks.append((label, e))

        mark("start")
        w.deposit(sense_enabled)
        mark("deposit")
        if getattr(w.b, "peer", False):
            # halo fused into the kernels (NVLink P2P stores): only an epoch flag crosses between ranks
            main = torch.cuda.current_stream()
            w.b.peer_signal()
            w.field_interior_a()
            mark("stencil_a")
            w.b.peer_wait()
            mark("wait_halo")
            w.agents(sense_enabled, move)
            mark("agents")
            if move:
                w.pack()
                mark("pack")
                self._comm(w.migrate_sends(), w.migrate_recvs(), self.ev_pack, self.ev_mig)
            w.field_rest()
            mark("stencil_b+edges")
            if move:
                main.wait_event(self.ev_mig)
                mark("wait_migration")
                w.unpack()
                mark("unpack")
            else:
                w.ticks += 1
            if marks is not None:
                self._marks.append(marks)
            return
        if self.overlap:
            main = torch.cuda.current_stream()
            self._comm(w.halo_sends(), w.halo_recvs(), self.ev_ready, self.ev_done)
            w.field_interior_a()                 # overlaps the halo
            mark("stencil_a")
            main.wait_event(self.ev_done)
            mark("wait_halo")
            w.agents(sense_enabled, move)
            mark("agents")
            if move:
                w.pack()
                mark("pack")
                self._comm(w.migrate_sends(), w.migrate_recvs(), self.ev_pack, self.ev_mig)
            w.field_rest()                       # overlaps the migration exchange
            mark("stencil_b+edges")
            if move:
                main.wait_event(self.ev_mig)
                mark("wait_migration")
                w.unpack()
                mark("unpack")
            else:
                w.ticks += 1
            if marks is not None:
                self._marks.append(marks)
            return
        for r in self._exchange(w.halo_sends(), w.halo_recvs()):
            r.wait()
        w.field_interior_a()
        w.agents(sense_enabled, move)
        if move:
            w.pack()
            for r in self._exchange(w.migrate_sends(), w.migrate_recvs()):
                r.wait()
        w.field_rest()
        if move:
            w.unpack()
        else:
            w.ticks += 1


class StripHostTick:
    """The host-facing tick of ONE strip (one process per GPU) — what `ReferenceCompat.tick_host` is on a single GPU:
    this rank's robots' poses come from pinned host memory ([3][slots]: x, y, theta in slot order), their wheel speeds
    and decisions go back to pinned host memory, the caller (Webots in the reference) owns the motion, so nothing
    moves or migrates on the device. The tick itself is `pf_strip_tick` in two calls: deposit, halo signal, first rows,
    halo wait, sense/decide (phases A and B), then the speeds start their way to the host on a side stream while the
    rest of the strip is stepped (phase C) — the supervisor needs the speeds, not the field."""

    def __init__(self, strip: "CudaStrip", worker: "StripWorker"):
        self.strip, self.w = strip, worker
        sw = strip.swarm
        n, dev = sw.n, sw.x.device
        self.d_pose = torch.empty((3, n), dtype=torch.float32, device=dev)
        self.d_dec = torch.empty(n, dtype=torch.uint8, device=dev)
        self.h_out = torch.empty((2, n), dtype=torch.float32).pin_memory()
        self.h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
        self.side = torch.cuda.Stream(dev)
        self.ev_b, self.ev_out = torch.cuda.Event(), torch.cuda.Event()
        self.h2d_bytes = self.d_pose.numel() * 4
        self.d2h_bytes = self.h_out.numel() * 4 + self.h_dec.numel()

    def tick(self, pose_host: torch.Tensor, sense_enabled: bool = True):
        strip, sw = self.strip, self.strip.swarm
        main = torch.cuda.current_stream(sw.x.device)
        self.d_pose.copy_(pose_host, non_blocking=True)
        if sense_enabled:
            strip.field.observe(sw, self.d_pose[0], self.d_pose[1], self.d_pose[2])
        else:  # before the start delay the reference stores no poses (sup:279-289 runs inside deposit_pheromone only)
            sw.x.copy_(self.d_pose[0]); sw.y.copy_(self.d_pose[1]); sw.theta.copy_(self.d_pose[2])
        flags = strip_flags(sense_enabled, False, False)
        strip.strip_tick(flags, abi.TICK_A | abi.TICK_B)
        self.ev_b.record(main)
        with torch.cuda.stream(self.side):      # the speeds leave beside the rest of the strip's stencil
            self.side.wait_event(self.ev_b)
            self.h_out[0].copy_(sw.wl, non_blocking=True)
            self.h_out[1].copy_(sw.wr, non_blocking=True)
            self.d_dec.copy_((sw.meta & abi.META_DEC_MASK) >> abi.META_DEC_SHIFT)
            self.h_dec.copy_(self.d_dec, non_blocking=True)
            self.ev_out.record(self.side)
        strip.strip_tick(flags, abi.TICK_C)
        self.w.ticks += 1
        main.wait_event(self.ev_out)            # the next tick's kernels rewrite wl / wr / meta
        self.ev_out.synchronize()
        return self.h_out[0], self.h_out[1], self.h_dec


def strip_flags(sense_enabled: bool, move: bool, hand_over: bool) -> int:
    """pf_strip_tick flags of a tick (deposit and sensing start together, sup:581)"""
    f = abi.STEP_FIELD
    if sense_enabled:
        f |= abi.STEP_DEPOSIT | abi.STEP_SENSE
    if move:
        f |= abi.STEP_MOVE
    if hand_over and sense_enabled and move:
        f |= abi.STEP_HANDOVER
    return f


def run_ticks(driver, nticks: int, sort_every: int = 0, sort=None, sense_enabled: bool = True,
              hand_over: bool = True, tick0: int = 0):
    """`nticks` ticks of a LocalDriver / DistDriver with the hand-over decided here: never on the last tick (the
    caller may read the field afterwards) nor on the tick before a re-sort. `sort()` re-orders the agents
    (strip.sort_agents on every strip) every `sort_every` ticks, counted from `tick0`."""
    for k in range(nticks):
        t = tick0 + k
        if sort is not None and sort_every > 0 and t % sort_every == 0:
            sort()
        before_sort = sort is not None and sort_every > 0 and (t + 1) % sort_every == 0
        driver.tick(sense_enabled, hand_over=hand_over and not before_sort and k + 1 < nticks)


def phase_means(driver: "DistDriver", skip: int = 3):
    """median milliseconds per phase over the profiled ticks of a DistDriver (after a synchronize);
    the median keeps one-off costs (NCCL connection set-up on the first exchanges) out"""
    acc = {}
    for marks in driver._marks[skip:]:
        for (_, e0), (label, e1) in zip(marks[:-1], marks[1:]):
            acc.setdefault(label, []).append(e0.elapsed_time(e1))
    return {k: float(np.median(v)) for k, v in acc.items()}


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
