"""numpy/ctypes face of oracle/pf_oracle.c — the float32 CPU restatement of the dense tick.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from swarm_robotics_pheromones_b200 import abi

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "libpf_oracle.so"
    src = _HERE / "pf_oracle.c"
    hdr = _HERE.parent / "include" / "pherofield.h"
    stale = (not so.exists()) or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime)
    if force or stale:
        cc = os.environ.get("CC", "gcc")
        subprocess.check_call(
            [cc, "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-shared",
             "-o", str(so), str(src), "-lm"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        fp = C.POINTER(C.c_float)
        u32p = C.POINTER(C.c_uint32)
        u8p = C.POINTER(C.c_uint8)
        u64p = C.POINTER(C.c_uint64)
        cfgp = C.POINTER(abi.PFConfig)
        L.po_sincosf.argtypes = [C.c_float, fp, fp]
        L.po_sincosf.restype = None
        L.po_atan2f.argtypes = [C.c_float, C.c_float]
        L.po_atan2f.restype = C.c_float
        L.po_pid_turn.argtypes = [cfgp, C.c_float, C.c_float, C.c_float, C.c_float]
        L.po_pid_turn.restype = C.c_float
        L.po_field_elems.argtypes = [cfgp]
        L.po_field_elems.restype = C.c_int64
        L.po_fill_ghosts.argtypes = [cfgp, fp]
        L.po_fill_ghosts.restype = None
        L.po_step_field.argtypes = [cfgp, fp, fp]
        L.po_step_field.restype = None
        L.po_deposit.argtypes = [cfgp, fp, C.c_int64, fp, fp, fp, u8p]
        L.po_deposit.restype = None
        for name in ("po_deposit_dense", "po_deposit_sorted"):
            getattr(L, name).argtypes = [cfgp, fp, C.c_int64, fp, fp, fp, u8p]
            getattr(L, name).restype = None
        L.po_agents_deposit.argtypes = [cfgp, fp, C.c_int64, fp, fp, u32p]
        L.po_agents_deposit.restype = C.c_int64
        L.po_sense.argtypes = [cfgp, fp, C.c_int64, fp, fp, fp, u32p, fp]
        L.po_sense.restype = None
        L.po_argmax5.argtypes = [fp]
        L.po_argmax5.restype = C.c_int
        L.po_agents_step.argtypes = [cfgp, fp, C.c_uint32, C.c_int64, fp, fp, fp, fp, fp, u32p, u8p,
                                     u64p]
        L.po_agents_step.restype = None
        L.po_tick.argtypes = [cfgp, fp, fp, C.POINTER(C.c_int), C.c_int, C.c_int64, fp, fp, fp, fp,
                              fp, u32p, u8p, u64p]
        L.po_tick.restype = None
        L.po_algo_dense_f64.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.po_algo_dense_f64.restype = C.c_double
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _u32(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32)) if a is not None else None


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def _u64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64)) if a is not None else None


def sincos(t: float):
    s = C.c_float()
    c = C.c_float()
    lib().po_sincosf(np.float32(t), C.byref(s), C.byref(c))
    return s.value, c.value


def atan2(y: float, x: float) -> float:
    return lib().po_atan2f(np.float32(y), np.float32(x))


def pid_turn(cfg, cx, cy, gx, gy) -> float:
    return lib().po_pid_turn(C.byref(cfg), cx, cy, gx, gy)


class AgentsNP:
    """Agents as numpy SoA (the oracle's copy of the pf_agents layout)."""

    FIELDS = ("x", "y", "theta", "wl", "wr", "meta", "gid")

    def __init__(self, x, y, theta, state=abi.STATE_SEARCHING, gid=None):
        n = len(x)
        self.x = np.ascontiguousarray(x, dtype=np.float32).copy()
        self.y = np.ascontiguousarray(y, dtype=np.float32).copy()
        self.theta = np.ascontiguousarray(theta, dtype=np.float32).copy()
        self.wl = np.zeros(n, np.float32)
        self.wr = np.zeros(n, np.float32)
        st = np.broadcast_to(np.asarray(state, dtype=np.uint32), (n,))
        self.meta = (st | abi.META_MOVED | (abi.DEC_NONE << abi.META_DEC_SHIFT)).astype(np.uint32)
        self.gid = (np.arange(n, dtype=np.uint32) if gid is None
                    else np.ascontiguousarray(gid, dtype=np.uint32).copy())

    @property
    def n(self):
        return len(self.x)

    def copy(self):
        o = object.__new__(AgentsNP)
        for f in self.FIELDS:
            setattr(o, f, getattr(self, f).copy())
        return o

    def select(self, idx):
        o = object.__new__(AgentsNP)
        for f in self.FIELDS:
            setattr(o, f, np.ascontiguousarray(getattr(self, f)[idx]))
        return o

    @staticmethod
    def concat(parts):
        o = object.__new__(AgentsNP)
        for f in AgentsNP.FIELDS:
            setattr(o, f, np.ascontiguousarray(np.concatenate([getattr(p, f) for p in parts])))
        return o


class DenseOracle:
    """Field buffers [C][rows+2][W] (ghost rows included) plus the tick functions."""

    def __init__(self, cfg: abi.PFConfig):
        self.cfg = cfg.copy()
        self.rows = cfg.row1 - cfg.row0
        n = lib().po_field_elems(C.byref(self.cfg))
        self.bufs = [np.zeros(n, np.float32), np.zeros(n, np.float32)]
        self.cur = 0
        self.counters = np.zeros(9, np.uint64)
        self.ticks = 0
        self.t0 = 0.0
        self.time = 0.0  # = t0 + ticks * double(cfg.dt), as pf_step does
        self.start_delay = 0.0

    # ---- views
    def _view(self, which):
        return self.bufs[which].reshape(self.cfg.C, self.rows + 2, self.cfg.W)

    @property
    def field(self):
        """owned rows of the current buffer, [C][rows][W] (a view)"""
        return self._view(self.cur)[:, 1:-1, :]

    @property
    def field_with_ghosts(self):
        return self._view(self.cur)

    def set_field(self, arr):
        self.field[...] = np.asarray(arr, np.float32).reshape(self.cfg.C, self.rows, self.cfg.W)

    # ---- pieces
    def fill_ghosts(self):
        lib().po_fill_ghosts(C.byref(self.cfg), _fp(self.bufs[self.cur]))

    def step_field(self, nsteps=1, fill_ghosts=True):
        for _ in range(nsteps):
            if fill_ghosts:
                self.fill_ghosts()
            lib().po_step_field(C.byref(self.cfg), _fp(self.bufs[self.cur]),
                                _fp(self.bufs[self.cur ^ 1]))
            self.cur ^= 1

    def deposit(self, x, y, amount, channel=None, impl="po_deposit"):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        amount = np.ascontiguousarray(amount, np.float32)
        ch = None if channel is None else np.ascontiguousarray(channel, np.uint8)
        getattr(lib(), impl)(C.byref(self.cfg), _fp(self.bufs[self.cur]), len(x), _fp(x), _fp(y),
                         _fp(amount), _u8(ch))

    def agents_deposit(self, a: AgentsNP) -> int:
        m = lib().po_agents_deposit(C.byref(self.cfg), _fp(self.bufs[self.cur]), a.n, _fp(a.x),
                                    _fp(a.y), _u32(a.meta))
        self.counters[8] += np.uint64(m)
        return m

    def sense(self, x, y, theta=None, meta=None):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        th = None if theta is None else np.ascontiguousarray(theta, np.float32)
        mt = None if meta is None else np.ascontiguousarray(meta, np.uint32)
        out = np.empty((len(x), 5), np.float32)
        lib().po_sense(C.byref(self.cfg), _fp(self.bufs[self.cur]), len(x), _fp(x), _fp(y), _fp(th),
                       _u32(mt), _fp(out))
        return out

    def agents_step(self, a: AgentsNP, flags=abi.STEP_SENSE | abi.STEP_MOVE):
        dec = np.empty(a.n, np.uint8)
        lib().po_agents_step(C.byref(self.cfg), _fp(self.bufs[self.cur]), flags, a.n, _fp(a.x),
                             _fp(a.y), _fp(a.theta), _fp(a.wl), _fp(a.wr), _u32(a.meta), _u8(dec),
                             _u64(self.counters))
        return dec

    def tick(self, a: AgentsNP, nsteps=1):
        """whole ticks on an unsharded field; returns the decisions of the last tick"""
        dec = np.empty(a.n, np.uint8)
        cur = C.c_int(self.cur)
        for _ in range(nsteps):
            sense = 1 if self.time >= self.start_delay else 0
            lib().po_tick(C.byref(self.cfg), _fp(self.bufs[0]), _fp(self.bufs[1]), C.byref(cur),
                          sense, a.n, _fp(a.x), _fp(a.y), _fp(a.theta), _fp(a.wl), _fp(a.wr),
                          _u32(a.meta), _u8(dec), _u64(self.counters))
            self.ticks += 1
            self.time = self.t0 + self.ticks * float(self.cfg.dt)
        self.cur = cur.value
        return dec


def cells_of(cfg: abi.PFConfig, x, y):
    """(row, col, inside) with the oracle's truncation rule, vectorised in float32."""
    x = np.asarray(x, np.float32)
    y = np.asarray(y, np.float32)
    inv = np.float32(1.0) / np.float32(cfg.cell_size)
    fr = (x - np.float32(cfg.origin_x)) * inv
    fc = (y - np.float32(cfg.origin_y)) * inv
    inside = (fr >= 0) & (fc >= 0) & (fr <= cfg.H) & (fc <= cfg.W)
    r = np.clip(np.trunc(fr), 0, cfg.H - 1).astype(np.int64)
    q = np.clip(np.trunc(fc), 0, cfg.W - 1).astype(np.int64)
    return r, q, inside


def algo_dense_f64(N: int, r: int, q: int, nticks: int):
    grid = np.zeros((N, N), np.float64)
    s = lib().po_algo_dense_f64(N, r, q, nticks, grid.ctypes.data_as(C.POINTER(C.c_double)))
    return s, grid
