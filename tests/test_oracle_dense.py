"""The dense float32 oracle (oracle/pf_oracle.c): pinned where the reference has arithmetic
(algo:351-356 closed form and golden grid, deposit rule, decision logic vs the trail oracle) and
cross-checked against independent implementations where it does not (diffusion: scipy correlate
with mode='reflect', mass conservation)."""
import json
import math

import numpy as np
import pytest

from oracle import dense, trail
from swarm_robotics_pheromones_b200 import abi


def test_algo_dense_closed_form_and_golden(golden_dir):
    g = json.load(open(golden_dir / "algo_dense.json"))
    s, grid = dense.algo_dense_f64(10, g["cell"][0], g["cell"][1], g["iterations"])
    assert abs(s - 9 * (1 - 0.9 ** g["iterations"])) < 1e-12
    assert np.array_equal(grid, np.array(g["grid"]))            # the real pheromone_algo.py output
    # the fp32 field path reproduces it to fp32 accuracy: deposit THEN evaporate, same tick
    cfg = abi.default_config(10, 10, 1, cell_size=1.0)
    o = dense.DenseOracle(cfg)
    for _ in range(g["iterations"]):
        o.deposit([g["cell"][0] + 0.5], [g["cell"][1] + 0.5], [1.0])
        o.step_field(1)
    ref = np.array(g["grid"])
    assert np.abs(o.field[0] - ref).max() <= 1e-6 * ref.max()  # fp32 vs f64, 25 compounded steps


def test_evap_keep_constant_is_float32_of_1_minus_0p1():
    cfg = abi.default_config(4, 4)
    assert np.float32(cfg.evap_keep[0]) == np.float32(1.0 - 0.1) == np.float32(0.9)


@pytest.mark.parametrize("stencil,kernel", [
    (5, np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float64)),
    (9, np.array([[1, 4, 1], [4, -20, 4], [1, 4, 1]], np.float64) / 6),
])
def test_diffusion_against_scipy_reflect(stencil, kernel):
    from scipy.ndimage import correlate
    rng = np.random.default_rng(0)
    cfg = abi.default_config(61, 47, 1)
    cfg.stencil = stencil
    cfg.diffusion[0] = 0.12
    cfg.evap_keep[0] = 1.0
    f = rng.random((61, 47)).astype(np.float32)
    o = dense.DenseOracle(cfg)
    o.set_field(f)
    o.step_field(12)
    g = f.astype(np.float64)
    for _ in range(12):
        g = g + np.float64(np.float32(0.12)) * correlate(g, kernel, mode="reflect")  # graph.py:24's mode
    assert np.abs(g - o.field[0]).max() < 2e-6
    assert abs(o.field.astype(np.float64).sum() - f.astype(np.float64).sum()) < 1e-3  # mass conserved


def test_graph_heatmap_properties(golden_dir):
    """graph.py:19-24 fixtures: deposit indexing [row = y, col = x] with abs(int()), +20 per point,
    mass conservation of the reflect-mode filter (sum = 760 * 20)."""
    g = json.load(open(golden_dir / "graph_heatmap.json"))
    assert g["shape"] == [497, 495] and g["n_points"] == 760
    assert abs(g["sum"] - 15200.0) < 1e-6
    assert g["argmax"] == [229, 152] and abs(g["max"] - 0.784932) < 1e-6
    # the dense oracle deposits the same cells when x/y are swapped into its (row = first coord) rule
    sup = json.load(open(golden_dir / "sup_replay_heatmap.json"))
    pts = np.array(sup["positions"])
    rows = np.abs((pts[:, 1] * 497).astype(int))
    cols = np.abs((pts[:, 0] * 495).astype(int))
    cfg = abi.default_config(497, 495, 1, cell_size=1.0)
    o = dense.DenseOracle(cfg)
    o.deposit(rows + 0.5, cols + 0.5, np.full(len(rows), 20.0))
    assert o.field.sum() == 15200.0
    ref = np.zeros((497, 495))
    np.add.at(ref, (rows, cols), 20)
    assert np.array_equal(o.field[0], ref.astype(np.float32))


def test_deterministic_math_close_to_libm():
    ts = np.linspace(-2 * math.pi, 2 * math.pi, 4001)
    err = max(max(abs(dense.sincos(t)[0] - math.sin(np.float32(t))), abs(dense.sincos(t)[1] - math.cos(np.float32(t))))
              for t in ts)
    assert err < 2e-7
    rng = np.random.default_rng(1)
    for _ in range(4000):
        y, x = rng.normal(size=2)
        assert abs(dense.atan2(y, x) - math.atan2(np.float32(y), np.float32(x))) < 5e-7
    assert dense.atan2(0.0, -1.0) == np.float32(math.pi)
    assert dense.atan2(1.0, 0.0) == np.float32(math.pi / 2)
    assert dense.atan2(0.0, 0.0) == 0.0


def test_pid_turn_matches_trail_oracle_to_fp32():
    cfg = abi.default_config(100, 100)
    rng = np.random.default_rng(2)
    for _ in range(2000):
        cx, cy, gx, gy = rng.uniform(-1, 1, 4)
        got = dense.pid_turn(cfg, cx, cy, gx, gy)
        want = trail.pid_turn((np.float32(cx), np.float32(cy)), (np.float32(gx), np.float32(gy)))
        d = abs(got - want)
        assert min(d, abs(d - 2 * math.pi)) < 5e-6


def test_ordered_argmax_first_maximum_wins():
    import ctypes as C
    L = dense.lib()

    def am(v):
        a = np.array(v, np.float32)
        return L.po_argmax5(a.ctypes.data_as(C.POINTER(C.c_float)))
    nan = float("nan")
    assert am([1, 1, 1, 1, 1]) == abi.DIR_CURRENT          # strict >, sup:447
    assert am([1, 2, 2, 2, 2]) == abi.DIR_FRONT
    assert am([0, 0, 0, 0, 5]) == abi.DIR_RIGHT
    assert am([nan, nan, 0, 3, 3]) == abi.DIR_LEFT         # absent directions are skipped
    assert am([nan] * 5) == -1


def test_deposit_rule_every_third_and_moved_gate():
    cfg = abi.default_config(32, 32, 1)
    o = dense.DenseOracle(cfg)
    a = dense.AgentsNP([0.155], [0.155], [0.0])
    amounts = []
    for k in range(9):
        before = float(o.field.sum())
        if k == 4:
            a.meta[0] &= ~np.uint32(abi.META_MOVED)   # not moved: "nothing deposited" (sup:321-322)
        o.agents_deposit(a)
        a.meta[0] |= np.uint32(abi.META_MOVED)
        amounts.append(float(o.field.sum()) - before)
    assert amounts == [1.0, 1.0, 20.0, 1.0, 0.0, 1.0, 20.0, 1.0, 1.0]
    assert abi.meta_count(int(a.meta[0])) == 8
    # homing: 5.0 and no every-3rd rule (sup:346-352)
    h = dense.AgentsNP([0.155], [0.155], [0.0], abi.STATE_HOMING)
    o2 = dense.DenseOracle(cfg)
    for _ in range(6):
        o2.agents_deposit(h)
    assert float(o2.field.sum()) == 30.0


def test_dense_decisions_match_trail_oracle_on_homing_golden(golden_dir):
    """The fp32 decide logic against the real reference's HOMING trace (through the pinned trail
    oracle): same decision wherever fp32 rounding does not sit on a round(.,2) boundary."""
    g = json.load(open(golden_dir / "sup_replay_homing.json"))
    cfg = abi.default_config(600, 600, 1, cell_size=0.005, ps=(g["ps"],) * 8)
    cfg.origin_x = cfg.origin_y = -1.5
    cfg.nest_x, cfg.nest_y = g["start"][0], g["start"][1]
    o = dense.DenseOracle(cfg)
    t = trail.TrailOracle(g["start"], g["food"], ps=(g["ps"],) * 8)
    t.state = trail.HOMING
    a = dense.AgentsNP([g["positions"][0][0]], [g["positions"][0][1]], [0.0], abi.STATE_HOMING)
    names = {"STRAIGHT": 0, "LEFT": 1, "RIGHT": 2, "KEEP": 3, "STOP": 4}
    mism = 0
    for rec, pos in zip(g["ticks"], g["positions"]):
        want = t.tick(rec["t"], pos)
        a.x[0], a.y[0] = pos[0], pos[1]
        d = int(o.agents_step(a, abi.STEP_SENSE)[0])
        if d != names[want["decision"]]:
            mism += 1
        else:
            assert abs(a.wl[0] - want["speeds"][0]) < 2e-5 * max(1, abs(want["speeds"][0]))
    assert mism == 0   # no round(.,2) boundary case on the shipped trace: fp32 decisions == the f64 reference's


def test_cells_truncate_toward_zero_and_clamp():
    cfg = abi.default_config(10, 10, 1, cell_size=0.1)
    r, q, inside = dense.cells_of(cfg, [0.0, 0.0999, 0.95, 1.0, -0.01, 1.2], [0.05] * 6)
    assert r.tolist() == [0, 0, 9, 9, 0, 9]
    assert inside.tolist() == [True, True, True, True, False, False]


def test_move_reflects_at_walls_and_sets_moved_flag():
    cfg = abi.default_config(100, 100, 1)
    o = dense.DenseOracle(cfg)
    a = dense.AgentsNP([0.0001, 0.5], [0.5, 0.5], [math.pi, 0.0], [abi.STATE_SEARCHING, abi.STATE_IDLE])
    o.agents_step(a)
    assert a.x[0] > 0.0 and abs(a.theta[0]) < 1e-3         # bounced off x = 0, heading mirrored
    assert abi.meta_moved(int(a.meta[0])) and not abi.meta_moved(int(a.meta[1]))  # idle: no motion
    # free-space speed: 2.572 rad/s * 0.02 m * 0.064 s = 3.29 mm per tick
    b = dense.AgentsNP([0.5], [0.5], [0.0])
    o.agents_step(b)
    assert abs((b.x[0] - 0.5) - 2.572 * 0.02 * 0.064) < 1e-6


def test_sorted_deposit_equals_dense_accumulator_bitexact():
    """po_deposit's two implementations (dense accumulator / stable sort by cell, used at the full-size configs)"""
    cfg = abi.default_config(300, 257, 3)
    rng = np.random.default_rng(0)
    n = 200000
    x = rng.uniform(-0.1, 3.1, n).astype(np.float32)
    y = rng.uniform(-0.1, 2.7, n).astype(np.float32)
    x[:5000] = 1.0                                  # one hot cell: a long in-order sum
    y[:5000] = 1.0
    amt = rng.uniform(0.1, 3, n).astype(np.float32)  # arbitrary amounts: the order of the sums matters
    ch = rng.integers(0, 4, n).astype(np.uint8)      # channel 3 does not exist: skipped
    a, b = dense.DenseOracle(cfg), dense.DenseOracle(cfg)
    a.set_field(rng.uniform(0, 1, (3, 300, 257)))
    b.set_field(a.field.copy())
    a.deposit(x, y, amt, ch, "po_deposit_dense")
    b.deposit(x, y, amt, ch, "po_deposit_sorted")
    assert (a.field.view(np.uint32) == b.field.view(np.uint32)).all()


def test_threaded_oracle_driver_equals_po_tick_bitexact():
    """tests/oracle_parallel.py (the oracle on several host threads, for the named-shape GPU tests) == po_tick"""
    from tests.oracle_parallel import ParallelOracle
    cfg = abi.default_config(96, 160, 2)
    cfg.stencil = 9
    cfg.fsm = 1
    for c in range(2):
        cfg.diffusion[c] = 0.1
    cfg.nest_x, cfg.nest_y = 0.48, 0.8
    cfg.food_x[0], cfg.food_y[0] = 0.7, 1.2
    cfg.food_radius = 0.1
    rng = np.random.default_rng(4)
    n = 3000
    x = rng.uniform(0.0, 0.96, n).astype(np.float32)
    y = rng.uniform(0.0, 1.6, n).astype(np.float32)
    th = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    st = rng.integers(1, 3, n)
    o, p = dense.DenseOracle(cfg), ParallelOracle(cfg, threads=3)
    a, b = dense.AgentsNP(x, y, th, st), dense.AgentsNP(x, y, th, st)
    for k in (1, 7, 30):
        d0 = o.tick(a, k)
        d1 = p.tick(b, k)
        assert (d0 == d1).all()
        for f in dense.AgentsNP.FIELDS:
            assert (getattr(a, f).view(np.uint32) == getattr(b, f).view(np.uint32)).all(), f
        assert (o.field.view(np.uint32) == p.field.view(np.uint32)).all()
        assert (o.counters == p.counters).all()
    p.close()
