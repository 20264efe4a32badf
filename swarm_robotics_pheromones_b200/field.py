"""Host-side face of the C-ABI library: PheroField (the double-buffered field + the hot-path
calls) and Swarm (agents, structure-of-arrays torch tensors as the buffer carrier).

PyTorch is plumbing here (device memory, streams); every compute call goes through ctypes into
libpherofield.so. There is no CPU path: constructing a PheroField without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _ffi, abi


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _DevView:
    """zero-copy torch view of library-owned device memory (via __cuda_array_interface__)"""

    def __init__(self, ptr: int, shape, strides_bytes, typestr="<f4"):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
            "strides": tuple(strides_bytes),
        }


class Swarm:
    """Agents SoA on one device. Mirrors pf_agents; slots with state EMPTY are holes."""

    FIELDS = ("x", "y", "theta", "wl", "wr", "meta", "gid")

    def __init__(self, x, y, theta, state=abi.STATE_SEARCHING, device="cuda", capacity=None,
                 gid=None):
        dev = torch.device(device)
        x = torch.as_tensor(x, dtype=torch.float32)
        n = x.numel()
        cap = n if capacity is None else int(capacity)
        if cap < n:
            raise ValueError("capacity smaller than the number of agents")
        self.device = dev

        def put(src, dtype):
            out = torch.zeros(cap, dtype=dtype, device=dev)
            out[:n] = torch.as_tensor(src, dtype=dtype).to(dev)
            return out

        self.x = put(x, torch.float32)
        self.y = put(y, torch.float32)
        self.theta = put(theta, torch.float32)
        self.wl = torch.zeros(cap, dtype=torch.float32, device=dev)
        self.wr = torch.zeros(cap, dtype=torch.float32, device=dev)
        st = np.broadcast_to(np.asarray(state, dtype=np.int64), (n,))
        meta = (st | abi.META_MOVED | (abi.DEC_NONE << abi.META_DEC_SHIFT)).astype(np.int64)
        # uint32 bit patterns carried in int32 tensors (torch has no full uint32 support)
        self.meta = put(meta.astype(np.uint32).view(np.int32), torch.int32)
        g = np.arange(n, dtype=np.uint32) if gid is None else np.asarray(gid, dtype=np.uint32)
        self.gid = put(g.view(np.int32), torch.int32)
        self._struct = None

    @property
    def n(self) -> int:
        return self.x.numel()

    def struct(self) -> abi.PFAgents:
        if self._struct is None:
            s = abi.PFAgents()
            s.n = self.n
            s.x, s.y, s.theta = self.x.data_ptr(), self.y.data_ptr(), self.theta.data_ptr()
            s.wl, s.wr = self.wl.data_ptr(), self.wr.data_ptr()
            s.meta, s.gid = self.meta.data_ptr(), self.gid.data_ptr()
            self._struct = s
        return self._struct

    # host copies as numpy with uint32 meta/gid
    def to_numpy(self) -> dict:
        out = {}
        for f in self.FIELDS:
            a = getattr(self, f).cpu().numpy()
            out[f] = a.view(np.uint32) if f in ("meta", "gid") else a
        return out

    def load_numpy(self, d: dict):
        for f in self.FIELDS:
            a = np.ascontiguousarray(d[f])
            if f in ("meta", "gid"):
                a = a.astype(np.uint32).view(np.int32)
            getattr(self, f).copy_(torch.from_numpy(a))

    def alive(self) -> torch.Tensor:
        return (self.meta & abi.META_STATE_MASK) != abi.STATE_EMPTY


class PheroField:
    """One strip (or the whole arena) of the pheromone field on one GPU."""

    def __init__(self, cfg: abi.PFConfig):
        if not torch.cuda.is_available():
            raise RuntimeError("pherofield needs a CUDA device (sm_100a); there is no CPU fallback")
        self._L = _ffi.lib()
        self.cfg = cfg.copy()
        self.device = torch.device("cuda", cfg.device)
        h = C.c_void_p()
        _ffi.check(self._L.pf_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        self.rows = cfg.row1 - cfg.row0
        self.start_delay = 0.0

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._L.pf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _s(self) -> int:
        return _stream_ptr(self.device)

    def _ck(self, rc):
        _ffi.check(rc, self._h)

    # -- field access
    def view(self, which: int = 0, with_ghosts: bool = False) -> torch.Tensor:
        """zero-copy torch view [C][rows(+2)][W] of the current (0) or next (1) buffer"""
        p = C.c_void_p()
        pitch = C.c_int64()
        self._ck(self._L.pf_field_ptr(self._h, which, 0, C.byref(p), C.byref(pitch)))
        pitch = pitch.value
        chan_stride = (self.rows + 2) * pitch
        if with_ghosts:
            base, rows = p.value - pitch * 4, self.rows + 2
        else:
            base, rows = p.value, self.rows
        v = _DevView(base, (self.cfg.C, rows, self.cfg.W), (chan_stride * 4, pitch * 4, 4))
        return torch.as_tensor(v, device=self.device)

    def fill(self, value: float, channel: int = -1):
        self._ck(self._L.pf_field_fill(self._h, channel, value, self._s()))

    def upload(self, arr, channel: Optional[int] = None):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if channel is None:
            a = a.reshape(self.cfg.C, self.rows, self.cfg.W)
            for c in range(self.cfg.C):
                self._ck(self._L.pf_field_upload_host(self._h, c, a[c].ctypes.data, self._s()))
        else:
            a = a.reshape(self.rows, self.cfg.W)
            self._ck(self._L.pf_field_upload_host(self._h, channel, a.ctypes.data, self._s()))
        torch.cuda.current_stream(self.device).synchronize()

    def snapshot(self, channel: Optional[int] = None) -> np.ndarray:
        if channel is None:
            out = np.empty((self.cfg.C, self.rows, self.cfg.W), np.float32)
            for c in range(self.cfg.C):
                self._ck(self._L.pf_snapshot_host(self._h, c, out[c].ctypes.data, self._s()))
            return out
        out = np.empty((self.rows, self.cfg.W), np.float32)
        self._ck(self._L.pf_snapshot_host(self._h, channel, out.ctypes.data, self._s()))
        return out

    def fill_ghosts(self):
        self._ck(self._L.pf_fill_ghosts(self._h, self._s()))

    def total(self, channel: int = 0) -> float:
        s = C.c_double()
        self._ck(self._L.pf_field_sum_host(self._h, channel, C.byref(s), self._s()))
        return s.value

    # -- hot path
    def reserve_agents(self, n: int):
        self._ck(self._L.pf_reserve_agents(self._h, n))

    def deposit(self, x: torch.Tensor, y: torch.Tensor, amount: torch.Tensor,
                channel: Optional[torch.Tensor] = None):
        self._ck(self._L.pf_deposit(self._h, x.numel(), _ptr(x), _ptr(y), _ptr(amount),
                                    _ptr(channel), self._s()))

    def agents_deposit(self, a: Swarm):
        self._ck(self._L.pf_agents_deposit(self._h, C.byref(a.struct()), self._s()))

    def step_field(self, nsteps: int = 1):
        self._ck(self._L.pf_step_field(self._h, nsteps, self._s()))

    def step_field_rows(self, r0: int, r1: int):
        self._ck(self._L.pf_step_field_rows(self._h, r0, r1, self._s()))

    def swap(self):
        self._ck(self._L.pf_swap(self._h))

    def sense(self, x: torch.Tensor, y: torch.Tensor, theta: Optional[torch.Tensor] = None,
              meta: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = torch.empty((x.numel(), 5), dtype=torch.float32, device=self.device)
        self._ck(self._L.pf_sense(self._h, x.numel(), _ptr(x), _ptr(y), _ptr(theta), _ptr(meta),
                                  _ptr(out), self._s()))
        return out

    def agents_step(self, a: Swarm, flags: int = abi.STEP_SENSE | abi.STEP_MOVE,
                    decisions: Optional[torch.Tensor] = None):
        self._ck(self._L.pf_agents_step(self._h, C.byref(a.struct()), flags, _ptr(decisions),
                                        self._s()))

    def step(self, a: Swarm, nsteps: int = 1, flags: int = abi.STEP_ALL):
        self._ck(self._L.pf_step_ex(self._h, C.byref(a.struct()), nsteps, flags, self._s()))

    def observe(self, a: Swarm, x: torch.Tensor, y: torch.Tensor, theta: Optional[torch.Tensor] = None):
        self._ck(self._L.pf_agents_observe(self._h, C.byref(a.struct()), _ptr(x), _ptr(y), _ptr(theta),
                                           self._s()))

    def sort_agents(self, a: Swarm):
        self._ck(self._L.pf_agents_sort(self._h, C.byref(a.struct()), self._s()))

    def set_time(self, time_s: float = 0.0, start_delay_s: float = 0.0):
        self._ck(self._L.pf_set_time(self._h, time_s, start_delay_s))
        self.start_delay = start_delay_s

    def time(self):
        t = C.c_double()
        k = C.c_uint64()
        self._ck(self._L.pf_get_time(self._h, C.byref(t), C.byref(k)))
        return t.value, k.value

    def evaporate_trail(self, elapsed: torch.Tensor, intensity: torch.Tensor, rate: float = 0.1):
        out = torch.empty_like(intensity)
        self._ck(self._L.pf_evaporate_trail(self._h, intensity.numel(), _ptr(elapsed),
                                            _ptr(intensity), rate, _ptr(out), self._s()))
        return out

    # -- migration
    def migrate_pack(self, a: Swarm, up: torch.Tensor, down: torch.Tensor, cap: int):
        self._ck(self._L.pf_migrate_pack(self._h, C.byref(a.struct()), _ptr(up), _ptr(down), cap,
                                         self._s()))

    def migrate_unpack(self, a: Swarm, buf: Optional[torch.Tensor], cap: int,
                       buf2: Optional[torch.Tensor] = None):
        """place the records of `buf`, then of `buf2`, into the lowest empty slots (one pass)"""
        self._ck(self._L.pf_migrate_unpack2(self._h, C.byref(a.struct()), _ptr(buf), _ptr(buf2), cap,
                                            self._s()))

    # -- peer-store halo (NVLink P2P)
    def peer_export(self) -> abi.PFPeerInfo:
        info = abi.PFPeerInfo()
        self._ck(self._L.pf_peer_export(self._h, C.byref(info)))
        return info

    def peer_ipc_handles(self) -> bytes:
        buf = (C.c_uint8 * 192)()
        self._ck(self._L.pf_peer_ipc_handles(self._h, buf))
        return bytes(buf)

    def peer_ipc_open(self, handles: bytes, dims: abi.PFPeerInfo) -> abi.PFPeerInfo:
        src = (C.c_uint8 * 192).from_buffer_copy(handles)
        out = abi.PFPeerInfo()
        self._ck(self._L.pf_peer_ipc_open(self._h, src, C.byref(dims), C.byref(out)))
        return out

    def peer_attach(self, up: Optional[abi.PFPeerInfo], down: Optional[abi.PFPeerInfo]):
        self._ck(self._L.pf_peer_attach(self._h, C.byref(up) if up is not None else None,
                                        C.byref(down) if down is not None else None))
        self._peer_keep = (up, down)

    def peer_push_edges(self):
        self._ck(self._L.pf_peer_push_edges(self._h, self._s()))

    def peer_signal(self):
        self._ck(self._L.pf_peer_signal(self._h, self._s()))

    def peer_wait(self):
        self._ck(self._L.pf_peer_wait(self._h, self._s()))

    def peer_status(self):
        """raises PFError(PF_ESTATE) if a bounded peer wait gave up (dead neighbour / strips out of lockstep)"""
        self._ck(self._L.pf_peer_status(self._h, self._s()))

    def strip_tick(self, a: Swarm, flags: int, phases: int = abi.TICK_ALL):
        """one whole strip tick by one C call (pf_strip_tick): leavers travel through the peer mailboxes"""
        self._ck(self._L.pf_strip_tick(self._h, C.byref(a.struct()), flags, phases, self._s()))

    STRIP_PHASES = ("deposit", "signal+stencil_a", "wait_halo", "agents", "pack", "stencil_rest", "wait_migration",
                    "unpack")

    def strip_profile(self, reset: bool = True) -> dict:
        """mean milliseconds per phase of pf_strip_tick over the ticks run since profile(True)"""
        ms = (C.c_double * 11)()
        n = C.c_uint64()
        self._ck(self._L.pf_strip_profile(self._h, ms, C.byref(n), 1 if reset else 0))
        k = max(1, n.value)
        return {name: ms[i] / k for i, name in enumerate(self.STRIP_PHASES)} | {"ticks": n.value}

    # -- counters / tuning
    def counters(self) -> abi.PFCounters:
        c = abi.PFCounters()
        self._ck(self._L.pf_counters_host(self._h, C.byref(c), self._s()))
        return c

    def reset_counters(self):
        self._ck(self._L.pf_counters_reset(self._h, self._s()))

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._ck(self._L.pf_launch_count(self._h, C.byref(n)))
        return n.value

    def tile_rows(self, swarm):
        """strips: local rows whose next-tick deposits the stencil can take (pf_tile_rows); (0, 0) = none"""
        r0, r1 = C.c_int(), C.c_int()
        self._ck(self._L.pf_tile_rows(self._h, C.byref(swarm.struct()), C.byref(r0), C.byref(r1)))
        return r0.value, r1.value

    def tile_stats(self):
        """(ticks whose stencil carried the next tick's deposits, how many of them fell back)"""
        a, f = C.c_uint64(), C.c_uint64()
        self._ck(self._L.pf_tile_stats(self._h, C.byref(a), C.byref(f)))
        return a.value, f.value

    def tile_dense(self):
        """(crowded tiles that deposit through a side buffer, side buffers available)"""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self._L.pf_tile_dense(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile(self, on: bool):
        self._ck(self._L.pf_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset: bool = True) -> abi.PFProfile:
        p = abi.PFProfile()
        self._ck(self._L.pf_profile_read(self._h, C.byref(p), 1 if reset else 0))
        return p

    def set_option(self, option: int, value: int):
        self._ck(self._L.pf_set_option(self._h, option, value))

    def set_stencil_variant(self, variant: int, rows_per_cta: int = 0):
        self._ck(self._L.pf_set_stencil_variant(self._h, variant, rows_per_cta))
