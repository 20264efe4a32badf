"""TrailSwarm — the reference's sparse timestamped trail (the shipped supervisor) for N robots on the
GPU, with the reference's method names.

Unlike the dense field, this engine reproduces the real reference's data structure: per robot a
dict {time: (intensity, position)} in insertion order, the age-window compounding evaporation
(sup:594-607), findNeighbours' radius/time filter (sup:366-376) and the every-3rd breadcrumb
(sup:316-318). One thread per robot; all robots tick in one kernel (`pf_trail_tick`).
`sup` = Phase-3/pheromone_algo_supervisor.py of the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _ffi, abi


class TrailSwarm:
    def __init__(self, robot_names: Sequence[str], start_positions, food_position, *, ps=60.0,
                 states=None, start_delay: float = 5.0, capacity: int = 64, device: int = 0,
                 record_sense: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("pherofield needs a CUDA device (sm_100a); there is no CPU fallback")
        self._L = _ffi.lib()
        self.names = list(robot_names)
        self.index: Dict[str, int] = {n: i for i, n in enumerate(self.names)}
        n = self.n = len(self.names)
        self.device = torch.device("cuda", device)
        self.start_delay = float(start_delay)
        h = C.c_void_p()
        rc = self._L.pf_trail_create(device, n, capacity, C.byref(h))
        if rc:
            raise _ffi.PFError(rc, (self._L.pf_trail_last_error(None) or b"").decode())
        self._h = h
        start = np.ascontiguousarray(np.asarray(start_positions, np.float64).reshape(n, 3))
        food = np.ascontiguousarray(np.asarray(food_position, np.float64).reshape(3))
        st = None
        if states is not None:
            st = np.ascontiguousarray(np.broadcast_to(np.asarray(states, np.uint8), (n,)))
        self._ck(self._L.pf_trail_configure(self._h, start.ctypes.data, food.ctypes.data,
                                            None if st is None else st.ctypes.data))
        dev = self.device
        self.pos = torch.zeros((n, 3), dtype=torch.float64, device=dev)
        self.ps = torch.full((n, 8), float(ps), dtype=torch.float64, device=dev)
        self.out = {
            "amount": torch.empty(n, dtype=torch.float64, device=dev),
            "n_global": torch.empty(n, dtype=torch.int32, device=dev),
            "direction": torch.empty(n, dtype=torch.int8, device=dev),
            "decision": torch.empty(n, dtype=torch.int8, device=dev),
            "wl": torch.empty(n, dtype=torch.float64, device=dev),
            "wr": torch.empty(n, dtype=torch.float64, device=dev),
            "n_live": torch.empty(n, dtype=torch.int32, device=dev),
            "sum_live": torch.empty(n, dtype=torch.float64, device=dev),
            "masks": torch.empty((5, n), dtype=torch.int64, device=dev),
        }
        if record_sense:  # snapshot of the entries the sense step saw (16 B x capacity per robot)
            self.out["pre_time"] = torch.empty((capacity, n), dtype=torch.float64, device=dev)
            self.out["pre_intensity"] = torch.empty((capacity, n), dtype=torch.float64, device=dev)
            self.out["n_pre"] = torch.empty(n, dtype=torch.int32, device=dev)
        o = abi.PFTrailOut()
        for k, t in self.out.items():
            setattr(o, k, t.data_ptr())
        self._out_struct = o
        self.time = 0.0

    def _ck(self, rc):
        if rc:
            raise _ffi.PFError(rc, (self._L.pf_trail_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.pf_trail_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, robot_name: str, homing: bool):
        self._ck(self._L.pf_trail_set_state(self._h, self.index[robot_name],
                                            abi.STATE_HOMING if homing else abi.STATE_SEARCHING))

    # ---- batched tick (device arrays in, device arrays out) -----------------------------------
    def tick_device(self, t_now: float, pos: Optional[torch.Tensor] = None):
        if pos is not None:
            self.pos.copy_(pos)
        self.time = float(t_now)
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._ck(self._L.pf_trail_tick(self._h, self.time, self.start_delay, self.pos.data_ptr(),
                                       self.ps.data_ptr(), C.byref(self._out_struct), s))
        return self.out

    def check_capacity(self):
        """raises PFError(PF_ESTATE) once a deposit did not fit a robot's live list (was silent)"""
        n = C.c_uint64()
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._ck(self._L.pf_trail_dropped(self._h, C.byref(n), s))

    # ---- live entries of one robot, as the reference's dict -----------------------------------
    def _entries(self, i: int):
        ptrs = [C.c_void_p() for _ in range(6)]
        cap = C.c_int()
        self._ck(self._L.pf_trail_entries(self._h, *[C.byref(p) for p in ptrs], C.byref(cap)))
        from .field import _DevView
        n, k = self.n, cap.value
        cols = []
        for p in ptrs[:5]:
            v = torch.as_tensor(_DevView(p.value, (k, n), (n * 8, 8), "<f8"), device=self.device)
            cols.append(v[:, i].cpu().numpy())
        nl = torch.as_tensor(_DevView(ptrs[5].value, (n,), (4,), "<i4"), device=self.device)
        m = int(nl[i].item())
        t, inten, x, y, z = (c[:m] for c in cols)
        return t, inten, np.stack([x, y, z], 1)

    def pheromone_grid(self, robot_name: str) -> dict:
        """{time: (intensity, [x, y, z])} in insertion order — pheromone_grids[name] (sup:237)"""
        t, inten, pos = self._entries(self.index[robot_name])
        return {float(tt): (float(ii), pp.tolist()) for tt, ii, pp in zip(t, inten, pos)}

    # ---- the reference's tick drivers ----------------------------------------------------------
    def startPheromone(self, robot_names=None, positions=None, t_now: Optional[float] = None):
        """Controller.startPheromone / homing (sup:565-607, 627-663) for every robot. `positions` is
        {name: [x, y, z]} or an [N, 3] array. Returns {name: record} with the values the reference
        computes mid-tick: amount, n_global, sense dict, direction, speeds, n_live, sum_live."""
        if positions is not None:
            if isinstance(positions, dict):
                p = self.pos.cpu().numpy()
                for name, q in positions.items():
                    p[self.index[name]] = np.asarray(q, np.float64)
            else:
                p = np.asarray(positions, np.float64).reshape(self.n, 3)
            self.pos.copy_(torch.from_numpy(np.ascontiguousarray(p)))
        if t_now is None:
            t_now = self.time
        names = self.names if robot_names is None else [nm for nm in robot_names if nm in self.index]
        out = self.tick_device(t_now)
        self.check_capacity()
        host = {k: v.cpu().numpy() for k, v in out.items()}
        res = {}
        for name in names:
            i = self.index[name]
            amount = host["amount"][i]
            rec = {
                "t": t_now,
                "amount": None if np.isnan(amount) else float(amount),
                "n_global": int(host["n_global"][i]),
                "direction": (None if host["direction"][i] < 0 else abi.DIR_NAMES[host["direction"][i]]),
                "decision": abi.DEC_NAMES[host["decision"][i]],
                "speeds": [float(host["wl"][i]), float(host["wr"][i])],
                "n_live": int(host["n_live"][i]),
                "sum_live": float(host["sum_live"][i]),
            }
            if "n_pre" in host and host["direction"][i] >= 0:
                # expand the device's membership masks into the reference's sensed dict (sup:390):
                # {'current': [[time, intensity], ...], 'front': ..., 'back': ..., 'left': ..., 'right': ...}
                m = int(host["n_pre"][i])
                tt, ii = host["pre_time"][:m, i], host["pre_intensity"][:m, i]
                sense = {}
                for d, dname in enumerate(abi.DIR_NAMES):
                    mask = int(np.uint64(host["masks"][d, i]))
                    sense[dname] = [[float(tt[k]), float(ii[k])] for k in range(m) if mask >> k & 1]
                rec["sense"] = sense
            else:
                rec["sense"] = None
            res[name] = rec
        return res

    homing = startPheromone
