"""CUDA vs the CPU oracle at the NAMED shapes of BASELINE.json configs 2, 3 and 4.

The oracle (oracle/pf_oracle.c) runs here on the box's host cores through tests/oracle_parallel.py: row blocks of
the stencil and slices of the agent arrays on several threads, every number still produced by the oracle's own C
functions. Compared bit for bit: every field cell, every agent record (position, heading, wheel speeds, state /
decision / deposit counter word), agent cell indices, the decision histogram and the deposit totals.

  config 2   8192 x 8192, 2^20 agents, 9-point, the first 100 ticks                 (north_star: bit-exact over 100 steps)
  config 3   32768 x 32768 x 2 channels, 2^24 agents, 9-point, 12 ticks at full size: channel 1 and the last rows of
             channel 0 lie beyond the 2^32-byte plane offset, 32 column blocks of the TMA stencil per row,
             the deposit riding on the stencil on 10 of the ticks
  config 4   16384 x 16384 x 2, 2^22 agents around the nest, nest + 8 food sources, FSM on, 12 ticks
Positions and decisions follow SURVEY §8a (sup:291-358, 379-417, 440-450; algo:351-356); the diffusion term and the
motion model are the oracle's definition (unpinned).
"""
import os

import numpy as np
import pytest
import torch

from oracle import dense
from swarm_robotics_pheromones_b200 import abi
from swarm_robotics_pheromones_b200.field import PheroField, Swarm
from tests.oracle_parallel import ParallelOracle

pytestmark = pytest.mark.gpu

THREADS = max(2, min(32, os.cpu_count() or 2))


def _agents(cfg, n, seed, cluster=None, two_states=False):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    xmax, ymax = cfg.H * cfg.cell_size, cfg.W * cfg.cell_size
    x = torch.rand(n, generator=g) * xmax
    y = torch.rand(n, generator=g) * ymax
    th = (torch.rand(n, generator=g) * 2 - 1) * float(np.pi)
    if cluster is not None:  # config 4: nest +- 5 % of the arena width
        x = cfg.nest_x + (torch.rand(n, generator=g) * 2 - 1) * cluster * xmax
        y = cfg.nest_y + (torch.rand(n, generator=g) * 2 - 1) * cluster * ymax
    st = torch.randint(1, 3, (n,), generator=g).numpy() if two_states else np.full(n, abi.STATE_SEARCHING)
    return x.float().numpy(), y.float().numpy(), th.float().numpy(), st


def _equal_bits_chunked(got, want, what, rows=2048):
    """bit comparison of two [C][H][W] float32 arrays without multi-GB temporaries"""
    assert got.shape == want.shape, (got.shape, want.shape)
    for c in range(got.shape[0]):
        for r0 in range(0, got.shape[1], rows):
            a = got[c, r0:r0 + rows].view(np.uint32)
            b = np.ascontiguousarray(want[c, r0:r0 + rows]).view(np.uint32)
            if not np.array_equal(a, b):
                bad = np.argwhere(a != b)
                r, q = bad[0]
                raise AssertionError(f"{what}: channel {c} row {r0 + r} col {q}: got {got[c, r0 + r, q]!r} want "
                                     f"{want[c, r0 + r, q]!r} ({len(bad)} cells differ in this block)")


def _compare(g, sw, o, a, what):
    got = sw.to_numpy()
    order = np.argsort(got["gid"], kind="stable")
    assert np.array_equal(got["gid"][order], a.gid), f"{what}: agent ids"
    for f in ("x", "y", "theta", "wl", "wr", "meta"):
        gv = got[f][order].view(np.uint32)
        wv = getattr(a, f).view(np.uint32)
        if not np.array_equal(gv, wv):
            i = int(np.flatnonzero(gv != wv)[0])
            raise AssertionError(f"{what}: agent field {f} differs first at gid {i}: got {got[f][order][i]!r} want "
                                 f"{getattr(a, f)[i]!r} ({int((gv != wv).sum())} agents)")
    # agent cell indices (trunc rule, algo:351) — implied by equal positions, stated because north_star names it
    r_o, q_o, _ = dense.cells_of(o.cfg, a.x, a.y)
    r_g, q_g, _ = dense.cells_of(o.cfg, got["x"][order], got["y"][order])
    assert np.array_equal(r_o, r_g) and np.array_equal(q_o, q_g), f"{what}: agent cell indices"
    c = g.counters()
    assert c.deposits == int(o.counters[8]), f"{what}: deposit totals {c.deposits} vs {int(o.counters[8])}"
    assert list(c.decisions) == [int(v) for v in o.counters[:6]], f"{what}: decision histogram"
    _equal_bits_chunked(g.snapshot(), o.field, f"{what}: field")


@pytest.mark.timeout(900)
def test_config2_full_size_first_100_ticks_bitexact_vs_oracle():
    cfg = abi.default_config(8192, 8192, 1)
    cfg.stencil = 9
    cfg.evap_keep[0] = 1.0 - 0.1
    cfg.diffusion[0] = 0.1
    n = 1 << 20
    x, y, th, st = _agents(cfg, n, seed=1234)
    o = ParallelOracle(cfg, THREADS)
    a = dense.AgentsNP(x, y, th, st)
    g = PheroField(cfg)
    sw = Swarm(x, y, th, st)
    done = 0
    for k in (1, 15, 16, 18, 50):          # re-sorted like the bench (every 16 ticks); compared five times
        g.sort_agents(sw)
        g.step(sw, k)
        o.tick(a, k)
        done += k
        _compare(g, sw, o, a, f"config 2 after {done} ticks")
    assert done == 100
    attempted, fell_back = g.tile_stats()
    assert attempted >= 90 and fell_back == 0   # the deposits rode on the stencil, as in the bench
    assert o.counters[abi.DEC_LEFT] + o.counters[abi.DEC_RIGHT] > 0
    g.close(); o.close()


@pytest.mark.timeout(1800)
def test_config3_full_size_12_ticks_bitexact_vs_oracle():
    free, _ = torch.cuda.mem_get_info()
    import psutil
    if free < 40 * 2**30 or psutil.virtual_memory().available < 48 * 2**30:
        pytest.skip("needs about 20 GB of device memory and 30 GB of host memory")
    H = W = 32768
    cfg = abi.default_config(H, W, 2)
    cfg.stencil = 9
    cfg.fsm = 1                       # both channels get deposits (SEARCHING -> 0, HOMING -> 1) and are sensed
    for c in range(2):
        cfg.evap_keep[c] = 1.0 - 0.1
        cfg.diffusion[c] = 0.1
    cfg.nest_x, cfg.nest_y = 0.5 * H * cfg.cell_size, 0.5 * W * cfg.cell_size
    cfg.food_x[0], cfg.food_y[0] = 0.7 * H * cfg.cell_size, 0.6 * W * cfg.cell_size
    cfg.food_radius = 2.0
    n = 1 << 24
    x, y, th, st = _agents(cfg, n, seed=1234, two_states=True)
    o = ParallelOracle(cfg, THREADS)
    a = dense.AgentsNP(x, y, th, st)
    g = PheroField(cfg)
    sw = Swarm(x, y, th, st)
    g.sort_agents(sw)
    for k in (3, 9):                  # two calls: the last tick of a call is never handed over
        g.step(sw, k)
        o.tick(a, k)
    torch.cuda.synchronize()
    attempted, fell_back = g.tile_stats()
    assert attempted == 2 + 8 and fell_back == 0
    _compare(g, sw, o, a, "config 3 after 12 ticks")
    # what this size is about: cells beyond the 2^32-byte offset of a buffer carry deposits and were compared
    f = o.field
    assert float(f[1, -64:, :].max()) > 0 and float(f[0, -64:, :].max()) > 0 and float(f[1, :64, :].max()) > 0
    assert o.counters[6] > 0          # some robots reached the food disc and switched trail
    g.close(); o.close()


@pytest.mark.timeout(900)
def test_config4_full_size_12_ticks_bitexact_vs_oracle():
    H = W = 16384
    cfg = abi.default_config(H, W, 2)
    cfg.stencil = 9
    cfg.fsm = 1
    for c in range(2):
        cfg.evap_keep[c] = 1.0 - 0.1
        cfg.diffusion[c] = 0.1
    cx, cy = 0.5 * H * cfg.cell_size, 0.5 * W * cfg.cell_size
    cfg.nest_x, cfg.nest_y = cx, cy
    cfg.n_food = 8
    for k in range(8):                # 8 food sources on a ring at 0.35 W
        cfg.food_x[k] = cx + 0.35 * W * cfg.cell_size * np.cos(2 * np.pi * k / 8)
        cfg.food_y[k] = cy + 0.35 * W * cfg.cell_size * np.sin(2 * np.pi * k / 8)
    cfg.food_radius = 0.16
    n = 1 << 22
    x, y, th, st = _agents(cfg, n, seed=1234, cluster=0.05, two_states=True)
    o = ParallelOracle(cfg, THREADS)
    a = dense.AgentsNP(x, y, th, st)
    g = PheroField(cfg)
    sw = Swarm(x, y, th, st)
    g.sort_agents(sw)
    for k in (4, 8):
        g.step(sw, k)
        o.tick(a, k)
        _compare(g, sw, o, a, f"config 4 after {k if k == 4 else 12} ticks")
    assert o.counters[abi.DEC_STOP] + o.counters[abi.DEC_KEEP] + o.counters[abi.DEC_LEFT] > 0   # HOMING branches near the nest
    g.close(); o.close()
