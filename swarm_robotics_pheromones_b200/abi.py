"""ctypes mirror of include/pherofield.h (structs and constants only; loads nothing).

Kept separate from `_ffi.py` so that host-side logic and the test oracle can share the struct
layout without touching the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math

PF_ABI_VERSION = 2
PF_MAX_CHANNELS = 4
PF_MAX_FOOD = 16
PF_MIGRATE_WORDS = 8

PF_OK, PF_EINVAL, PF_ECUDA, PF_ENOMEM, PF_ESTATE = 0, -1, -2, -3, -4

STATE_EMPTY, STATE_SEARCHING, STATE_HOMING, STATE_IDLE = 0, 1, 2, 3
DIR_CURRENT, DIR_FRONT, DIR_BACK, DIR_LEFT, DIR_RIGHT = 0, 1, 2, 3, 4
DIR_NAMES = ("current", "front", "back", "left", "right")  # dict order of sup:390
DEC_STRAIGHT, DEC_LEFT, DEC_RIGHT, DEC_KEEP, DEC_STOP, DEC_NONE = 0, 1, 2, 3, 4, 5
DEC_NAMES = ("STRAIGHT", "LEFT", "RIGHT", "KEEP", "STOP", "NONE")
SENSE_WORLD, SENSE_HEADING = 0, 1

META_STATE_MASK = 0x3
META_MOVED = 0x4
META_DEC_SHIFT = 3
META_DEC_MASK = 0x38
META_CARRY = 0x40
META_HASPOS = 0x80
META_COUNT_SHIFT = 8
META_COUNT_MASK = 0x7FFFFF
META_HANDED = 0x80000000

OPT_OVERLAP_AGENTS = 1
OPT_NO_FUSED_KEYS = 2
OPT_USE_GRAPH = 3
OPT_NO_OWNER_DEPOSIT = 4
OPT_FORCE_OWNER_DEPOSIT = 5
OPT_NO_TILE_DEPOSIT = 6
OPT_PEER_TIMEOUT_MS = 7
OPT_STRIP_ONE_LAUNCH = 8
OPT_STRIP_PART_A_TILES = 9
OPT_STRIP_MIG_INLINE = 11
OPT_STRIP_NO_PREDEPOSIT = 12
OPT_STRIP_EDGES_SIDE = 13
TICK_A, TICK_B, TICK_C, TICK_ALL = 0x1, 0x2, 0x4, 0x7
STEP_DEPOSIT, STEP_SENSE, STEP_FIELD, STEP_MOVE, STEP_ALL = 0x1, 0x2, 0x4, 0x8, 0xF
STEP_HANDOVER = 0x10  # pf_agents_step on a strip: interior rows hand their next deposit to the stencil

# reference constants --------------------------------------------------------------------------
# Braitenberg coefficients, Phase-3/pheromone_algo_supervisor.py:77
BRAITENBERG_COEFFICIENTS = (
    (0.942, -0.22), (0.63, -0.1), (0.5, -0.06), (-0.06, -0.06),
    (-0.06, -0.06), (-0.06, 0.5), (-0.19, 0.63), (-0.13, 0.942),
)
MAX_SPEED = 3.14  # sup:74


def braitenberg_sums(ps=(0.0,) * 8):
    """sum_j coeff[j][i] * (1 - ps_j / (1024/2)) accumulated in the order of sup:468-470."""
    left = 0.0
    right = 0.0
    for j in range(8):
        left += BRAITENBERG_COEFFICIENTS[j][0] * (1.0 - (ps[j] / (1024 / 2)))
        right += BRAITENBERG_COEFFICIENTS[j][1] * (1.0 - (ps[j] / (1024 / 2)))
    return left, right


class PFConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("row0", C.c_int32), ("row1", C.c_int32),
        ("cell_size", C.c_float),
        ("origin_x", C.c_float),
        ("origin_y", C.c_float),
        ("stencil", C.c_int32),
        ("evap_keep", C.c_float * PF_MAX_CHANNELS),
        ("diffusion", C.c_float * PF_MAX_CHANNELS),
        ("delete_threshold", C.c_float),
        ("intensity_explore", C.c_float),
        ("intensity_high", C.c_float),
        ("intensity_homing", C.c_float),
        ("high_every", C.c_int32),
        ("move_threshold", C.c_float),
        ("deposit_channel", C.c_int32 * 4),
        ("sense_channel", C.c_int32 * 4),
        ("sense_mode", C.c_int32),
        ("sense_drow", C.c_int32 * 5),
        ("sense_dcol", C.c_int32 * 5),
        ("sense_dist", C.c_float),
        ("pid_kp", C.c_float), ("pid_kd", C.c_float),
        ("cruise_speed", C.c_float),
        ("max_speed", C.c_float),
        ("braitenberg_l", C.c_float),
        ("braitenberg_r", C.c_float),
        ("home_radius", C.c_float),
        ("home_lo_cm", C.c_float), ("home_hi_cm", C.c_float),
        ("nest_x", C.c_float), ("nest_y", C.c_float),
        ("n_food", C.c_int32),
        ("food_x", C.c_float * PF_MAX_FOOD),
        ("food_y", C.c_float * PF_MAX_FOOD),
        ("food_radius", C.c_float),
        ("fsm", C.c_int32),
        ("dt", C.c_float),
        ("wheel_radius", C.c_float),
        ("axle_length", C.c_float),
        ("migrate_cap", C.c_int32),
    ]

    def copy(self) -> "PFConfig":
        out = PFConfig()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(PFConfig))
        return out

    def as_dict(self) -> dict:
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class PFAgents(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("x", C.c_void_p), ("y", C.c_void_p), ("theta", C.c_void_p),
        ("wl", C.c_void_p), ("wr", C.c_void_p),
        ("meta", C.c_void_p), ("gid", C.c_void_p),
    ]


class PFCounters(C.Structure):
    _fields_ = [
        ("ticks", C.c_uint64),
        ("deposits", C.c_uint64),
        ("decisions", C.c_uint64 * 6),
        ("food_grabbed", C.c_uint64),
        ("food_delivered", C.c_uint64),
        ("migrated_out", C.c_uint64),
        ("migrated_in", C.c_uint64),
        ("migrate_dropped", C.c_uint64),
        ("peer_timeouts", C.c_uint64),
    ]


class PFTrailOut(C.Structure):
    _fields_ = [("amount", C.c_void_p), ("n_global", C.c_void_p), ("direction", C.c_void_p),
                ("decision", C.c_void_p), ("wl", C.c_void_p), ("wr", C.c_void_p),
                ("n_live", C.c_void_p), ("sum_live", C.c_void_p), ("masks", C.c_void_p),
                ("pre_time", C.c_void_p), ("pre_intensity", C.c_void_p), ("n_pre", C.c_void_p)]


class PFPeerInfo(C.Structure):
    _fields_ = [("buf0", C.c_void_p), ("buf1", C.c_void_p), ("flags", C.c_void_p),
                ("rows", C.c_int64), ("pitch", C.c_int64), ("chan_stride", C.c_int64),
                ("mig", C.c_void_p), ("mig_cap", C.c_int64)]


class PFProfile(C.Structure):
    _fields_ = [("ticks", C.c_uint64), ("deposit_ms", C.c_double), ("stencil_ms", C.c_double),
                ("agents_ms", C.c_double)]


def default_config(H: int, W: int, C_: int = 1, *, cell_size: float = 0.01, device: int = 0,
                   ps=(0.0,) * 8) -> PFConfig:
    """Python mirror of pf_config_default() (csrc/pherofield.cu); tests check they agree.

    Constants: sup:230-250 (intensities, rates), sup:74-77 (MAX_SPEED, coefficients),
    sup:260 (move threshold), sup:437 (PID gains), sup:489,667 (home thresholds),
    algo:247-249 (dense grid intensity / evaporation rate).
    """
    c = PFConfig()
    c.abi_version = PF_ABI_VERSION
    c.device = device
    c.H, c.W, c.C = H, W, C_
    c.row0, c.row1 = 0, H
    c.cell_size = cell_size
    c.origin_x = 0.0
    c.origin_y = 0.0
    c.stencil = 0
    for k in range(PF_MAX_CHANNELS):
        c.evap_keep[k] = 1.0 - 0.1
        c.diffusion[k] = 0.0
    c.delete_threshold = 0.0
    c.intensity_explore = 1.0
    c.intensity_high = 20.0
    c.intensity_homing = 5.0
    c.high_every = 3
    c.move_threshold = 0.0015
    for s in range(4):
        c.deposit_channel[s] = 0
        c.sense_channel[s] = 0
    if C_ >= 2:  # two-trail foraging: searchers lay the home trail and follow the food trail
        c.deposit_channel[STATE_SEARCHING] = 0
        c.deposit_channel[STATE_HOMING] = 1
        c.sense_channel[STATE_SEARCHING] = 1
        c.sense_channel[STATE_HOMING] = 0
    c.sense_mode = SENSE_HEADING
    drow = (0, 0, 0, -1, 1)
    dcol = (0, 1, -1, 0, 0)
    for d in range(5):
        c.sense_drow[d] = drow[d]
        c.sense_dcol[d] = dcol[d]
    c.sense_dist = cell_size
    c.pid_kp = 1.0
    c.pid_kd = 0.01
    c.cruise_speed = 1.0
    c.max_speed = MAX_SPEED
    bl, br = braitenberg_sums(ps)
    c.braitenberg_l = bl
    c.braitenberg_r = br
    c.home_radius = 0.7
    c.home_lo_cm = 5.0
    c.home_hi_cm = 6.0
    c.nest_x = 0.5 * H * cell_size
    c.nest_y = 0.5 * W * cell_size
    c.n_food = 1
    c.food_x[0] = 0.75 * H * cell_size
    c.food_y[0] = 0.75 * W * cell_size
    c.food_radius = 0.05
    c.fsm = 0
    c.dt = 0.064
    c.wheel_radius = 0.02
    c.axle_length = 0.052
    c.migrate_cap = 0
    return c


def pack_meta(state: int, moved: bool = True, count: int = 0, dec: int = DEC_NONE) -> int:
    return (state & META_STATE_MASK) | (META_MOVED if moved else 0) | (dec << META_DEC_SHIFT) | (
        count << META_COUNT_SHIFT)


def meta_state(m):
    return m & META_STATE_MASK


def meta_count(m):
    return (m >> META_COUNT_SHIFT) & META_COUNT_MASK


def meta_decision(m):
    return (m & META_DEC_MASK) >> META_DEC_SHIFT


def meta_moved(m):
    return (m & META_MOVED) != 0


def ceil_log2(n: int) -> int:
    return max(1, math.ceil(math.log2(max(2, n))))
