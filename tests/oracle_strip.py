"""Test-only backend for swarm_robotics_pheromones_b200.sharded: the CPU oracle as the per-strip
compute engine, numpy memory shared with CPU torch tensors so gloo can move the bytes."""
import numpy as np
import torch

from oracle import dense
from swarm_robotics_pheromones_b200 import abi

W8 = abi.PF_MIGRATE_WORDS


class OracleStrip:
    """hand_over=True also emulates the library's PF_STEP_HANDOVER protocol (pf_tile_rows): the robots that
    end up in the interior rows have their NEXT deposit added to those rows of the next buffer by the
    stencil call over exactly that range, and are skipped by the next deposit() — so that the driver's
    ordering of the phases can be checked on CPU."""

    def __init__(self, cfg, x, y, theta, state, gid, capacity, migrate_cap, hand_over=False):
        self.cfg = cfg
        self.hand_over = hand_over
        self._delta = None      # (r0, r1, [C][r1-r0][W] sums) waiting for the stencil call over (r0, r1)
        self._handed = None     # (r0, r1) whose robots the next deposit() skips
        self.handed_ticks = 0
        self.o = dense.DenseOracle(cfg)
        self.rows = cfg.row1 - cfg.row0
        n = len(x)
        pad = capacity - n
        z = np.zeros(pad, np.float32)
        self.a = dense.AgentsNP(np.concatenate([x, z]), np.concatenate([y, z]),
                                np.concatenate([theta, z]),
                                np.concatenate([np.asarray(state, np.int64), np.zeros(pad, np.int64)]),
                                np.concatenate([np.asarray(gid, np.uint32), np.zeros(pad, np.uint32)]))
        self.a.meta[n:] = abi.STATE_EMPTY
        self.mcap = migrate_cap
        words = W8 * (migrate_cap + 1)
        self.send_up = torch.zeros(words, dtype=torch.int32)
        self.send_down = torch.zeros(words, dtype=torch.int32)
        self.recv_up = torch.zeros(words, dtype=torch.int32)
        self.recv_down = torch.zeros(words, dtype=torch.int32)

    def _rows(self, idx):
        v = self.o.field_with_ghosts
        return [torch.from_numpy(v[c, idx]) for c in range(self.cfg.C)]

    def edge_rows(self):
        return self._rows(1), self._rows(self.rows)

    def ghost_rows(self):
        return self._rows(0), self._rows(self.rows + 1)

    def tile_rows(self):
        """the library's rule in small: whole 64-row bands, not the first one, not the last 64+ rows next to a
        neighbour"""
        if not self.hand_over:
            return (0, 0)
        lo = 64
        hi = ((self.rows - 64) // 64) * 64 if self.cfg.row1 < self.cfg.H else self.rows
        return (lo, hi) if hi > lo else (0, 0)

    def _depositors_in(self, r0, r1):
        a = self.a
        r, _, inside = dense.cells_of(self.cfg, a.x, a.y)
        st = a.meta & 3
        moved = (a.meta & abi.META_MOVED) != 0
        lr = r - self.cfg.row0
        return ((st == abi.STATE_SEARCHING) | (st == abi.STATE_HOMING)) & moved & inside & (lr >= r0) & (lr < r1)

    def _deposit_subset(self, oracle, idx):
        sub = self.a.select(idx)
        m = oracle.agents_deposit(sub)
        self.a.meta[idx] = sub.meta
        return m

    def deposit(self):
        if self._handed is None:
            self.o.agents_deposit(self.a)
            return
        r0, r1 = self._handed
        self._handed = None
        rest = np.nonzero(~self._depositors_in(r0, r1))[0]   # array order is kept
        self._deposit_subset(self.o, rest)

    def field_rows(self, r0, r1):
        # the oracle steps whole strips: do it once, on the first call of the tick
        if not getattr(self, "_stepped", False):
            import ctypes as C
            self.o.fill_ghosts()  # reflect rows at the arena's outer edges only
            dense.lib().po_step_field(C.byref(self.o.cfg), dense._fp(self.o.bufs[self.o.cur]),
                                      dense._fp(self.o.bufs[self.o.cur ^ 1]))
            self._stepped = True
        if self._delta is not None and (r0, r1) == self._delta[:2]:
            nxt = self.o._view(self.o.cur ^ 1)[:, 1:-1, :]
            nxt[:, r0:r1, :] += self._delta[2]            # next[cell] = stencil value + ordered sum
            self._handed = (r0, r1)
            self._delta = None
            self.handed_ticks += 1

    def agents(self, flags):
        self.o.agents_step(self.a, flags & ~abi.STEP_HANDOVER)
        if flags & abi.STEP_HANDOVER:
            r0, r1 = self.tile_rows()
            assert r1 > r0, "hand-over asked for without a row range"
            idx = np.nonzero(self._depositors_in(r0, r1))[0]
            if not hasattr(self, "_o2"):
                self._o2 = dense.DenseOracle(self.cfg)
            self._o2.field[...] = 0
            m = self._deposit_subset(self._o2, idx)       # bumps their counters now, as the move kernel does
            self.o.counters[8] += np.uint64(m)
            self._delta = (r0, r1, self._o2.field[:, r0:r1, :].copy())

    def swap(self):
        assert self._delta is None, "handed-over rows were not stepped as one call"
        self.o.cur ^= 1
        self._stepped = False

    def pack(self):
        a = self.a
        r, _, _ = dense.cells_of(self.cfg, a.x, a.y)
        alive = (a.meta & 3) != 0
        for buf, mask in ((self.send_up, alive & (r < self.cfg.row0)),
                          (self.send_down, alive & (r >= self.cfg.row1))):
            idx = np.nonzero(mask)[0][: self.mcap]
            out = buf.numpy().view(np.uint32)
            out[:] = 0
            out[0] = len(idx)
            rec = out[W8:].reshape(-1, W8)
            for k, f in enumerate(("x", "y", "theta", "wl", "wr")):
                rec[: len(idx), k] = getattr(a, f)[idx].view(np.uint32)
            rec[: len(idx), 5] = a.meta[idx]
            rec[: len(idx), 6] = a.gid[idx]
            a.meta[idx] = abi.STATE_EMPTY

    def unpack(self, from_up, from_down):
        """arrivals from the upper neighbour take the lowest empty slots (ascending), arrivals from the
        lower neighbour the highest ones (descending) — the library's placement rule"""
        a = self.a
        for use, buf, from_top in ((from_up, self.recv_up, True), (from_down, self.recv_down, False)):
            if not use:
                continue
            empty = np.nonzero((a.meta & 3) == 0)[0]
            inp = buf.numpy().view(np.uint32)
            m = int(inp[0])
            rec = inp[W8:].reshape(-1, W8)[:m]
            slots = empty[:m] if from_top else empty[::-1][:m]
            assert len(slots) == m, "no empty slot left"
            for k, f in enumerate(("x", "y", "theta", "wl", "wr")):
                getattr(a, f)[slots] = rec[:, k].view(np.float32)
            a.meta[slots] = rec[:, 5]
            a.gid[slots] = rec[:, 6]

    def synchronize(self):
        pass


def gather_agents_sorted(parts):
    """concatenate alive agents of several AgentsNP-like dicts and sort by gid"""
    cat = {f: np.concatenate([p[f] for p in parts]) for f in dense.AgentsNP.FIELDS}
    alive = (cat["meta"] & 3) != 0
    cat = {f: v[alive] for f, v in cat.items()}
    order = np.argsort(cat["gid"], kind="stable")
    return {f: v[order] for f, v in cat.items()}
