"""Test-only: the CPU oracle (oracle/pf_oracle.c) driven by several host threads, so that the BASELINE configs
can be checked at their named sizes (8192^2 ... 32768^2 x 2) in seconds instead of minutes.

Nothing is restated here: every number still comes from the oracle's own C functions.
  * evaporate+diffuse: po_step_field on row blocks of ONE shared pair of buffers. A block [a, b) of channel ch is
    handed to the oracle as a one-channel strip whose buffer starts at the block's upper neighbour row, so its
    "ghost rows" are simply the adjacent rows of the big buffer (the arena's outer ghosts are filled first,
    po_fill_ghosts). Blocks write disjoint rows of the next buffer.
  * sense/decide/move: po_agents_step on disjoint slices of the agent arrays (an agent's update reads only its own
    record and the field), per-thread counters added up afterwards.
  * deposit: po_agents_deposit as is (one thread; at these sizes it takes the stable-sort path of po_deposit).
ctypes releases the GIL during the calls. tests/test_oracle_dense.py checks this driver against DenseOracle.tick.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from oracle import dense
from swarm_robotics_pheromones_b200 import abi


class ParallelOracle:
    def __init__(self, cfg: abi.PFConfig, threads: int = 8):
        assert cfg.row0 == 0 and cfg.row1 == cfg.H, "whole arenas only"
        self.o = dense.DenseOracle(cfg)
        self.cfg = self.o.cfg
        self.threads = max(1, int(threads))
        self.pool = ThreadPoolExecutor(self.threads)
        H, W = cfg.H, cfg.W
        nblk = min(H, self.threads * 4)
        edges = np.linspace(0, H, nblk + 1).astype(np.int64)
        self.blocks = []  # (sub-config, float offset into a buffer)
        plane = (H + 2) * W
        for ch in range(cfg.C):
            for a, b in zip(edges[:-1], edges[1:]):
                if b <= a:
                    continue
                sub = cfg.copy()
                sub.C = 1
                sub.row0, sub.row1 = int(a), int(b)
                sub.evap_keep[0] = cfg.evap_keep[ch]
                sub.diffusion[0] = cfg.diffusion[ch]
                self.blocks.append((sub, ch * plane + int(a) * W))

    # same attributes the tests read on DenseOracle
    @property
    def field(self):
        return self.o.field

    @property
    def counters(self):
        return self.o.counters

    def _step_field(self):
        o = self.o
        o.fill_ghosts()
        cur, nxt = o.bufs[o.cur], o.bufs[o.cur ^ 1]
        L = dense.lib()
        fp = C.POINTER(C.c_float)

        def one(blk):
            sub, off = blk
            L.po_step_field(C.byref(sub), C.cast(cur.ctypes.data + 4 * off, fp), C.cast(nxt.ctypes.data + 4 * off, fp))

        list(self.pool.map(one, self.blocks))
        o.cur ^= 1

    def _agents_step(self, a: dense.AgentsNP, flags: int, dec: np.ndarray):
        o = self.o
        L = dense.lib()
        n = a.n
        T = self.threads
        cuts = np.linspace(0, n, T + 1).astype(np.int64)
        counters = np.zeros((T, 9), np.uint64)
        cur = o.bufs[o.cur]

        def one(t):
            i0, i1 = int(cuts[t]), int(cuts[t + 1])
            if i1 <= i0:
                return
            L.po_agents_step(C.byref(o.cfg), dense._fp(cur), flags, i1 - i0, dense._fp(a.x[i0:i1]),
                             dense._fp(a.y[i0:i1]), dense._fp(a.theta[i0:i1]), dense._fp(a.wl[i0:i1]),
                             dense._fp(a.wr[i0:i1]), dense._u32(a.meta[i0:i1]), dense._u8(dec[i0:i1]),
                             dense._u64(counters[t]))

        list(self.pool.map(one, range(T)))
        o.counters += counters.sum(axis=0)

    def tick(self, a: dense.AgentsNP, nsteps: int = 1):
        """po_tick's order (sup:575-607, 753-754): deposit -> sense/decide(+move) -> evaporate+diffuse"""
        o = self.o
        dec = np.empty(a.n, np.uint8)
        for _ in range(nsteps):
            sense = o.time >= o.start_delay
            flags = abi.STEP_MOVE | (abi.STEP_SENSE if sense else 0)
            if sense:
                o.agents_deposit(a)
            self._agents_step(a, flags, dec)
            self._step_field()
            o.ticks += 1
            o.time = o.t0 + o.ticks * float(o.cfg.dt)
        return dec

    def close(self):
        self.pool.shutdown()
