"""Sense riding on the stencil ("pre-sense", pf_deposit_tile.cuh / PF_OPT_PRESENSE): the stencil CTA that has just
written a tile (and applied the next tick's deposits to it) reads the five cells each robot on it will sense next
tick out of L2 and leaves the ordered arg-max (sup:440-450) in one byte per robot; the next move kernel uses it
instead of gathering from HBM. Same cells, same values, same comparison — so everything must stay bit-identical
to the CPU oracle, with the option on, off, under CUDA-graph replay, on ragged tiles, in crowded tiles whose
request bin overflows, and for sensing geometries the tile cannot serve."""
import numpy as np
import pytest

from oracle import dense
from swarm_robotics_pheromones_b200 import abi
from swarm_robotics_pheromones_b200.field import PheroField, Swarm
from tests.test_gpu_parity import assert_bitexact, make_cfg, rand_agents

pytestmark = pytest.mark.gpu


def _foraging_cfg(H, W, mode, stencil=9):
    cfg = make_cfg(H, W, 2, stencil)
    cfg.sense_mode = mode
    cfg.fsm = 1
    cfg.nest_x, cfg.nest_y = 0.8, 7.0
    cfg.food_x[0], cfg.food_y[0] = 1.1, 9.0
    cfg.food_radius = 0.3
    return cfg


def _by_gid(sw, a_np, what):
    got = sw.to_numpy()
    order = np.argsort(got["gid"], kind="stable")
    for f in ("x", "y", "theta", "wl", "wr"):
        assert_bitexact(got[f][order], getattr(a_np, f), f"{what}: {f}")
    assert (got["meta"][order] == a_np.meta).all(), f"{what}: meta"


def _observe_np(a, x, y, th):
    """pf_agents_observe on the oracle's arrays: storingPosition + has_moved (sup:260-289), fp32 op for op"""
    dx, dy = x - a.x, y - a.y
    moved = np.sqrt(dx * dx + dy * dy, dtype=np.float32) >= np.float32(0.0015)
    moved |= (a.meta & np.uint32(abi.META_HASPOS)) == 0
    a.x[:], a.y[:], a.theta[:] = x, y, th
    a.meta = ((a.meta & ~np.uint32(abi.META_MOVED)) | np.where(moved, abi.META_MOVED, 0).astype(np.uint32)
              | np.uint32(abi.META_HASPOS))


@pytest.mark.parametrize("graph", [0, 1])
@pytest.mark.parametrize("mode", [abi.SENSE_WORLD, abi.SENSE_HEADING])
def test_sense_riding_on_the_stencil_bitexact(mode, graph):
    # 3 x 2 tiles per channel, the last ones ragged (160 = 64 + 64 + 32 rows, 1536 = 1024 + 512 columns); a random
    # field so that all three SEARCHING decisions occur; SEARCHING robots sense channel 1, HOMING ones deposit there
    cfg = _foraging_cfg(160, 1536, mode)
    rng = np.random.default_rng(5)
    f0 = rng.random((2, 160, 1536), dtype=np.float32)
    n = 8000   # about 1070 depositors / requests per full tile and channel (lists take 1408, request bins 1280)
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.0)
    x[:40] = 0.0                                  # wall huggers: absent directions, never served
    y[40:80] = cfg.W * cfg.cell_size
    st = rng.integers(1, 3, n)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    fields = []
    for on in (1, 0):
        g = PheroField(cfg)
        g.upload(f0)
        g.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
        g.set_option(abi.OPT_USE_GRAPH, graph)
        g.set_option(abi.OPT_PRESENSE, on)
        sw = Swarm(x, y, th, st)
        g.sort_agents(sw)
        fields.append((g, sw))
    consumed_ticks = 0
    searching_min = n
    for k in (1, 2, 9, 40):
        o.tick(a_np, k)
        consumed_ticks += k - 1   # the first tick of a call gathers; the last one asks for nothing
        searching_min = min(searching_min, int(((a_np.meta & 3) == abi.STATE_SEARCHING).sum()))
        for on, (g, sw) in zip((1, 0), fields):
            g.step(sw, k)
            assert_bitexact(g.snapshot(), o.field, f"presense={on}: field")
            _by_gid(sw, a_np, f"presense={on}")
            c = g.counters()
            assert list(c.decisions) == [int(v) for v in o.counters[:6]], f"presense={on}: decision histogram"
    assert o.counters[abi.DEC_LEFT] > 0 and o.counters[abi.DEC_RIGHT] > 0 and o.counters[abi.DEC_STRAIGHT] > 0
    served_on, served_off = fields[0][0].presense_stats(), fields[1][0].presense_stats()
    assert served_off == 0
    # centre cells off the tile borders: (62 + 62 + 30) / 160 of the rows, (1022 + 510) / 1536 of the columns
    assert served_on > 0.7 * (154 / 160) * (1532 / 1536) * searching_min * consumed_ticks, served_on
    assert fields[0][0].tile_stats()[1] == 0
    for g, _ in fields:
        g.close()


def test_sense_riding_not_served_when_the_probes_leave_the_3x3_block():
    # heading-relative probes 2.5 cells out: the request format (one 3 x 3 block around the centre) cannot express
    # them, so nobody is served and every robot gathers — results unchanged
    cfg = _foraging_cfg(128, 1024, abi.SENSE_HEADING)
    cfg.sense_dist = 2.5 * cfg.cell_size
    rng = np.random.default_rng(6)
    f0 = rng.random((2, 128, 1024), dtype=np.float32)
    n = 5000
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.05)
    st = rng.integers(1, 3, n)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    g = PheroField(cfg)
    g.upload(f0)
    g.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    g.set_option(abi.OPT_PRESENSE, 1)
    sw = Swarm(x, y, th, st)
    g.sort_agents(sw)
    o.tick(a_np, 12)
    g.step(sw, 12)
    assert_bitexact(g.snapshot(), o.field, "field")
    _by_gid(sw, a_np, "far probes")
    assert g.presense_stats() == 0
    g.close()


def test_sense_riding_crowded_tile_overflowing_request_bin():
    # 5000 SEARCHING robots inside one tile (request bin: 1280): the first 1280 requests are served, the others
    # gather; 1500 robots on a 4 x 4 patch overflow the tile's deposit bins as well, so the first hand-over falls
    # back (that tick nobody is served: the cells do not hold the deposits yet) and the tile then goes dense
    cfg = make_cfg(128, 2048, 1, 9)
    rng = np.random.default_rng(78)
    f0 = rng.random((1, 128, 2048), dtype=np.float32)
    n = 9000
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.02)
    x[:5000] = (0.10 + rng.uniform(0, 0.40, 5000)).astype(np.float32)
    y[:5000] = (12.0 + rng.uniform(0, 4.00, 5000)).astype(np.float32)
    x[:1500] = (0.30 + rng.uniform(0, 0.04, 1500)).astype(np.float32)
    y[:1500] = (14.0 + rng.uniform(0, 0.04, 1500)).astype(np.float32)
    st = np.full(n, abi.STATE_SEARCHING)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    g = PheroField(cfg)
    g.upload(f0)
    g.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    g.set_option(abi.OPT_PRESENSE, 1)
    g.set_option(abi.OPT_USE_GRAPH, 0)
    sw = Swarm(x, y, th, st)
    g.sort_agents(sw)
    for k in (3, 20):
        o.tick(a_np, k)
        g.step(sw, k)
        assert_bitexact(g.snapshot(), o.field, "field with a crowded tile")
        _by_gid(sw, a_np, "crowded tile")
    attempted, fell_back = g.tile_stats()
    assert attempted == 2 + 19 and fell_back == 1
    served = g.presense_stats()
    # 20 served ticks (21 hand-overs, one fell back); per tick at most 1280 in the crowded tile + the 4000 elsewhere
    assert 20 * 1280 < served <= 20 * (1280 + 4000), served
    assert list(g.counters().decisions) == [int(v) for v in o.counters[:6]]
    g.close()


def test_sense_riding_survives_pose_replacement_between_calls():
    # bytes never outlive the move kernel behind the stencil that wrote them: replacing the poses between two
    # pf_step calls (pf_agents_observe) and re-sorting must leave no stale direction behind
    cfg = _foraging_cfg(128, 1024, abi.SENSE_WORLD)
    rng = np.random.default_rng(9)
    f0 = rng.random((2, 128, 1024), dtype=np.float32)
    n = 6000
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.05)
    st = np.full(n, abi.STATE_SEARCHING)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    g = PheroField(cfg)
    g.upload(f0)
    g.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    g.set_option(abi.OPT_PRESENSE, 1)
    sw = Swarm(x, y, th, st)
    g.sort_agents(sw)
    import torch
    for rnd in range(3):
        o.tick(a_np, 5)
        g.step(sw, 5)
        _by_gid(sw, a_np, f"round {rnd}")
        # new poses for everybody, in gid order on the oracle side and in array order on the device
        nx = rng.uniform(0.05, cfg.H * cfg.cell_size - 0.05, n).astype(np.float32)
        ny = rng.uniform(0.05, cfg.W * cfg.cell_size - 0.05, n).astype(np.float32)
        nth = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
        _observe_np(a_np, nx, ny, nth)
        gid = sw.gid.cpu().numpy().astype(np.int64)
        g.observe(sw, torch.from_numpy(nx[gid]).cuda(), torch.from_numpy(ny[gid]).cuda(), torch.from_numpy(nth[gid]).cuda())
        if rnd == 1:
            g.sort_agents(sw)
    o.tick(a_np, 4)
    g.step(sw, 4)
    assert_bitexact(g.snapshot(), o.field, "field")
    _by_gid(sw, a_np, "after pose replacement")
    assert g.presense_stats() > 0
    g.close()
