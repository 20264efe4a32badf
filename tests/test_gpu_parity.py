"""GPU parity: the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded
inputs. Integer/index work and (because op order is fixed and FMA-free) the fp32 field values are
compared BIT-EXACT; the north_star tolerance for field values (1e-5 of max after 1000 steps) is
asserted as well."""
import numpy as np
import pytest
import torch

from oracle import dense
from swarm_robotics_pheromones_b200 import abi
from swarm_robotics_pheromones_b200.field import PheroField, Swarm

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bitexact(a, b, what=""):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, what
    bad = bits(a) != bits(b)
    # +0 / -0 never arise on this path; NaNs compare by bit pattern class
    bad &= ~(np.isnan(a) & np.isnan(b))
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} differ; first at {np.argwhere(bad)[:3].tolist()} " \
                          f"gpu={a[bad][:3]} oracle={b[bad][:3]}"


def make_cfg(H, W, C=1, stencil=5, D=0.1, **kw):
    cfg = abi.default_config(H, W, C)
    cfg.stencil = stencil
    for k in range(C):
        cfg.diffusion[k] = D * (1.0 + 0.5 * k)
        cfg.evap_keep[k] = 0.9 + 0.03 * k
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def rand_agents(rng, cfg, n, margin=0.0, states=None):
    xmax = cfg.H * cfg.cell_size
    ymax = cfg.W * cfg.cell_size
    x = rng.uniform(margin, xmax - margin, n).astype(np.float32)
    y = rng.uniform(margin, ymax - margin, n).astype(np.float32)
    th = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    st = abi.STATE_SEARCHING if states is None else states
    return x, y, th, st


# ----------------------------------------------------------------------------- K1 stencil
@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("stencil", [0, 5, 9])
@pytest.mark.parametrize("shape", [(64, 128, 1), (37, 256, 2), (130, 1024, 1), (96, 2176, 2), (33, 132, 1)])
def test_stencil_bitexact(shape, stencil, variant):
    H, W, C = shape
    rng = np.random.default_rng(H * 1000 + W + stencil)
    cfg = make_cfg(H, W, C, stencil)
    f0 = rng.random((C, H, W), dtype=np.float32) * 10
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    g = PheroField(cfg)
    g.set_stencil_variant(variant, 16)
    g.upload(f0)
    o.step_field(7)
    g.step_field(7)
    assert_bitexact(g.snapshot(), o.field, f"stencil {stencil} {shape} v{variant}")
    g.close()


@pytest.mark.parametrize("stencil", [0, 5, 9])
def test_stencil_odd_width_scalar_path(stencil):
    # 497 x 495 is the heat-map of reference graph.py:10
    cfg = make_cfg(497, 495, 1, stencil)
    rng = np.random.default_rng(5)
    f0 = rng.random((1, 497, 495), dtype=np.float32)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    g = PheroField(cfg)
    g.upload(f0)
    o.step_field(5)
    g.step_field(5)
    assert_bitexact(g.snapshot(), o.field)
    g.close()


def test_stencil_delete_threshold():
    cfg = make_cfg(64, 256, 1, 5)
    cfg.delete_threshold = 0.1  # sup:603
    rng = np.random.default_rng(11)
    f0 = (rng.random((1, 64, 256), dtype=np.float32) * 0.3).astype(np.float32)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    g = PheroField(cfg)
    g.upload(f0)
    o.step_field(6)
    g.step_field(6)
    assert (o.field == 0).any()
    assert_bitexact(g.snapshot(), o.field)
    g.close()


def test_stencil_mass_conservation_large():
    # size-independent property at a size the oracle would not finish quickly: with keep = 1 the
    # reflect-boundary Laplacian conserves the total (graph.py: sum stays 15200)
    H = W = 4096
    cfg = make_cfg(H, W, 1, 9, D=0.15)
    cfg.evap_keep[0] = 1.0
    g = PheroField(cfg)
    v = g.view()
    torch.manual_seed(0)
    v.copy_(torch.rand((1, H, W), device="cuda"))
    s0 = g.total(0)
    g.step_field(20)
    s1 = g.total(0)
    assert abs(s1 - s0) / s0 < 1e-6
    # and pure evaporation scales the total by keep^n
    cfg2 = make_cfg(H, W, 1, 0)
    cfg2.evap_keep[0] = 0.9
    g2 = PheroField(cfg2)
    g2.view().copy_(g.view())
    g2.step_field(3)
    assert abs(g2.total(0) / s1 - 0.9 ** 3) < 1e-5
    g.close()
    g2.close()


# ----------------------------------------------------------------------------- K2 deposit
@pytest.mark.parametrize("C", [1, 2])
def test_deposit_generic_bitexact(C):
    cfg = make_cfg(96, 160, C, 5)
    rng = np.random.default_rng(21)
    n = 20000
    x = rng.uniform(-0.05, cfg.H * cfg.cell_size + 0.05, n).astype(np.float32)  # some outside
    y = rng.uniform(-0.05, cfg.W * cfg.cell_size + 0.05, n).astype(np.float32)
    # hot cells: many agents in one cell, long runs crossing warp boundaries
    x[:3000] = 0.335
    y[:3000] = 0.777
    x[3000:3100] = 0.0
    y[3000:3100] = 0.0
    amount = rng.random(n, dtype=np.float32) * 3  # arbitrary fp32: order matters
    ch = rng.integers(0, C, n).astype(np.uint8)
    f0 = rng.random((C, 96, 160), dtype=np.float32)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    o.deposit(x, y, amount, ch)
    g = PheroField(cfg)
    g.upload(f0)
    g.deposit(torch.tensor(x).cuda(), torch.tensor(y).cuda(), torch.tensor(amount).cuda(),
              torch.tensor(ch).cuda())
    assert_bitexact(g.snapshot(), o.field, "generic deposit")
    g.close()


def test_deposit_empty_and_all_outside():
    cfg = make_cfg(32, 64, 1, 0)
    g = PheroField(cfg)
    e = torch.empty(0, device="cuda")
    g.deposit(e, e, e)
    x = torch.full((100,), -1.0, device="cuda")
    g.deposit(x, x, torch.ones(100, device="cuda"))
    assert g.total(0) == 0.0
    g.close()


def test_agents_deposit_rule():
    # every-3rd rule on the agent's own counter (sup:316-318), homing amount, moved gate
    cfg = make_cfg(64, 64, 2, 0)
    rng = np.random.default_rng(3)
    n = 5000
    x, y, th, _ = rand_agents(rng, cfg, n)
    st = rng.integers(1, 4, n)
    a_np = dense.AgentsNP(x, y, th, st)
    a_np.meta[::7] &= ~np.uint32(abi.META_MOVED)
    a_np.meta |= (rng.integers(0, 50, n).astype(np.uint32) << abi.META_COUNT_SHIFT)
    o = dense.DenseOracle(cfg)
    g = PheroField(cfg)
    sw = Swarm(x, y, th, st)
    sw.load_numpy({f: getattr(a_np, f) for f in a_np.FIELDS})
    for _ in range(4):
        o.agents_deposit(a_np)
        g.agents_deposit(sw)
    got = sw.to_numpy()
    assert (got["meta"] == a_np.meta).all()
    assert_bitexact(g.snapshot(), o.field)
    assert g.counters().deposits == int(o.counters[8])
    g.close()


# ----------------------------------------------------------------------------- K3 sense/decide/move
@pytest.mark.parametrize("mode", [abi.SENSE_WORLD, abi.SENSE_HEADING])
def test_sense_bitexact(mode):
    cfg = make_cfg(80, 96, 2, 5)
    cfg.sense_mode = mode
    rng = np.random.default_rng(8)
    f0 = rng.random((2, 80, 96), dtype=np.float32)
    n = 4000
    x, y, th, _ = rand_agents(rng, cfg, n)
    x[:50] = 0.0  # wall-hugging agents: absent directions
    y[50:100] = cfg.W * cfg.cell_size
    st = rng.integers(1, 3, n)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    g = PheroField(cfg)
    g.upload(f0)
    sw = Swarm(x, y, th, st)
    want = o.sense(a_np.x, a_np.y, a_np.theta, a_np.meta)
    got = g.sense(sw.x, sw.y, sw.theta, sw.meta).cpu().numpy()
    assert np.isnan(want).any()
    assert (np.isnan(want) == np.isnan(got)).all()
    assert_bitexact(got, want)
    g.close()


def _compare_agents(sw, a_np, what=""):
    got = sw.to_numpy()
    for f in ("meta", "gid"):
        assert (got[f] == getattr(a_np, f)).all(), f"{what}: {f} differs"
    for f in ("x", "y", "theta", "wl", "wr"):
        assert_bitexact(got[f], getattr(a_np, f), f"{what}: {f}")


@pytest.mark.parametrize("mode", [abi.SENSE_WORLD, abi.SENSE_HEADING])
def test_agents_step_bitexact(mode):
    cfg = make_cfg(128, 128, 1, 5)
    cfg.sense_mode = mode
    cfg.nest_x, cfg.nest_y = 0.64, 0.64
    rng = np.random.default_rng(17)
    f0 = rng.random((1, 128, 128), dtype=np.float32)
    n = 30000
    x, y, th, _ = rand_agents(rng, cfg, n)
    st = rng.integers(0, 4, n)  # includes EMPTY and IDLE
    # homing agents around the nest at every distance band (STOP ring, inside radius, far)
    x[:3000] = (0.64 + rng.uniform(-0.08, 0.08, 3000)).astype(np.float32)
    y[:3000] = (0.64 + rng.uniform(-0.08, 0.08, 3000)).astype(np.float32)
    st[:3000] = abi.STATE_HOMING
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    g = PheroField(cfg)
    g.upload(f0)
    sw = Swarm(x, y, th, st)
    dec_g = torch.empty(n, dtype=torch.uint8, device="cuda")
    for it in range(5):
        dec_o = o.agents_step(a_np)
        g.agents_step(sw, decisions=dec_g)
        assert (dec_g.cpu().numpy() == dec_o).all(), f"decisions differ at iteration {it}"
        _compare_agents(sw, a_np, f"iteration {it}")
    hist = np.bincount(dec_o, minlength=6)
    assert hist[abi.DEC_STOP] > 0 and hist[abi.DEC_KEEP] > 0 and hist[abi.DEC_LEFT] > 0
    c = g.counters()
    assert list(c.decisions) == [int(v) for v in o.counters[:6]]
    g.close()


# ----------------------------------------------------------------------------- whole tick
def _run_ticks(cfg, n_agents, nticks, seed, check_every, start_delay=0.0, states=None, margin=0.05):
    rng = np.random.default_rng(seed)
    x, y, th, st = rand_agents(rng, cfg, n_agents, margin=margin, states=states)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    o.start_delay = start_delay
    g = PheroField(cfg)
    g.set_time(0.0, start_delay)
    sw = Swarm(x, y, th, st)
    done = 0
    while done < nticks:
        k = min(check_every, nticks - done)
        o.tick(a_np, k)
        g.step(sw, k)
        done += k
        _compare_agents(sw, a_np, f"tick {done}")
        r_o, q_o, _ = dense.cells_of(cfg, a_np.x, a_np.y)
        got = sw.to_numpy()
        r_g, q_g, _ = dense.cells_of(cfg, got["x"], got["y"])
        assert (r_o == r_g).all() and (q_o == q_g).all(), "agent cell indices differ"
        assert_bitexact(g.snapshot(), o.field, f"field at tick {done}")
    c = g.counters()
    assert c.deposits == int(o.counters[8]), "deposit totals differ"
    assert list(c.decisions) == [int(v) for v in o.counters[:6]], "decision histogram differs"
    assert c.ticks == nticks
    return o, g, a_np, sw


def test_tick_config1_first_100_steps_bitexact():
    # BASELINE config 1: 1024 x 1024 single channel, 4096 agents, evaporate+diffuse+deposit
    cfg = make_cfg(1024, 1024, 1, 5, D=0.1)
    cfg.evap_keep[0] = 0.9
    o, g, a_np, sw = _run_ticks(cfg, 4096, 100, seed=1234, check_every=10)
    assert o.counters[abi.DEC_LEFT] + o.counters[abi.DEC_RIGHT] > 0
    g.close()


def test_tick_field_after_1000_steps_within_tolerance():
    cfg = make_cfg(256, 256, 1, 9, D=0.1)
    cfg.evap_keep[0] = 0.98
    o, g, a_np, sw = _run_ticks(cfg, 2048, 1000, seed=99, check_every=250)
    got, want = g.snapshot(), o.field
    tol = 1e-5 * float(want.max())  # north_star: within 1e-5 of the maximum level
    assert np.abs(got - want).max() <= tol
    g.close()


def test_tick_two_channel_fsm_foraging():
    # config-4-like: two trails, nest + 8 food sources on a ring, FSM on
    cfg = make_cfg(256, 256, 2, 5, D=0.05)
    cfg.fsm = 1
    cfg.nest_x, cfg.nest_y = 1.28, 1.28
    cfg.n_food = 8
    for k in range(8):
        cfg.food_x[k] = 1.28 + 0.35 * 2.56 * np.cos(2 * np.pi * k / 8)
        cfg.food_y[k] = 1.28 + 0.35 * 2.56 * np.sin(2 * np.pi * k / 8)
    cfg.food_radius = 0.1
    rng = np.random.default_rng(5)
    n = 4000
    x = (1.28 + rng.uniform(-0.9, 0.9, n)).astype(np.float32)
    y = (1.28 + rng.uniform(-0.9, 0.9, n)).astype(np.float32)
    th = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    st = rng.integers(1, 3, n)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    g = PheroField(cfg)
    sw = Swarm(x, y, th, st)
    for _ in range(6):
        o.tick(a_np, 50)
        g.step(sw, 50)
        _compare_agents(sw, a_np)
        assert_bitexact(g.snapshot(), o.field)
    c = g.counters()
    assert c.food_grabbed == int(o.counters[6]) and c.food_delivered == int(o.counters[7])
    assert c.food_grabbed > 0 and c.food_delivered > 0
    g.close()


def test_tick_overlapped_agent_stream_same_result():
    cfg = make_cfg(256, 1024, 2, 5)
    rng = np.random.default_rng(77)
    x, y, th, _ = rand_agents(rng, cfg, 5000, margin=0.05)
    a_np = dense.AgentsNP(x, y, th)
    o = dense.DenseOracle(cfg)
    g = PheroField(cfg)
    g.set_option(abi.OPT_OVERLAP_AGENTS, 1)
    sw = Swarm(x, y, th)
    o.tick(a_np, 40)
    g.step(sw, 40)
    _compare_agents(sw, a_np)
    assert_bitexact(g.snapshot(), o.field)
    g.close()


def test_tick_start_delay_random_walk():
    # before t = 5 s no deposit and no sensing (sup:731-744); 100 ticks cross the 5 s mark at 79
    cfg = make_cfg(128, 256, 1, 5)
    o, g, a_np, sw = _run_ticks(cfg, 1000, 100, seed=4, check_every=20, start_delay=5.0)
    assert int(o.counters[abi.DEC_NONE]) == 1000 * 79
    g.close()


# ----------------------------------------------------------------------------- migration
def test_migrate_pack_unpack_roundtrip():
    H, W = 64, 64
    cfg = make_cfg(H, W, 1, 5)
    top = cfg.copy(); top.row0, top.row1 = 0, 32
    bot = cfg.copy(); bot.row0, bot.row1 = 32, 64
    gt, gb = PheroField(top), PheroField(bot)
    rng = np.random.default_rng(2)
    n, cap_slots, cap = 1000, 1500, 256
    x, y, th, st = rand_agents(rng, cfg, n)
    x[:] = rng.uniform(0.25, 0.3199, n)  # all in the top strip, near the boundary (row 32 = 0.32)
    x[:100] = rng.uniform(0.32, 0.33, 100)  # these have crossed
    st_arr = np.ones(n, np.int64)
    sw_t = Swarm(x, y, th, st_arr, capacity=cap_slots)
    sw_b = Swarm(np.zeros(0), np.zeros(0), np.zeros(0), capacity=cap_slots)
    words = abi.PF_MIGRATE_WORDS * (cap + 1)
    up = torch.zeros(words, dtype=torch.int32, device="cuda")
    down = torch.zeros(words, dtype=torch.int32, device="cuda")
    gt.migrate_pack(sw_t, up, down, cap)
    assert int(up[0]) == 0 and int(down[0]) == 100
    gb.migrate_unpack(sw_b, down, cap)
    a_t, a_b = sw_t.to_numpy(), sw_b.to_numpy()
    alive_t = (a_t["meta"] & 3) != 0
    alive_b = (a_b["meta"] & 3) != 0
    assert alive_t.sum() == 900 and alive_b.sum() == 100
    assert sorted(a_b["gid"][alive_b].tolist()) == list(range(100))
    # deterministic placement: k-th record -> k-th empty slot, records ordered by source slot
    assert a_b["gid"][:100].tolist() == list(range(100))
    assert_bitexact(a_b["x"][:100], x[:100])
    c = gt.counters()
    assert c.migrated_out == 100 and gb.counters().migrated_in == 100
    gt.close(); gb.close()


def test_evaporate_trail_known_answers(golden_dir):
    # sup:519-525 compounded once per tick while 2 <= age <= 4 reproduces pheromonelevels.txt
    import json
    lv = json.load(open(golden_dir / "pheromonelevels.json"))["e2"]
    cfg = make_cfg(32, 32, 1, 0)
    g = PheroField(cfg)
    ages = torch.tensor([2.048 + 0.064 * k for k in range(14)], dtype=torch.float64, device="cuda")
    inten = torch.tensor([20.0], dtype=torch.float64, device="cuda")
    vals = []
    for k in range(14):
        inten = g.evaporate_trail(ages[k:k + 1], inten)
        vals.append(float(inten))
    for want, nticks in zip(lv[:4], (14, 13, 12, 11)):
        assert abs(vals[nticks - 1] - want) < 1e-12
    g.close()


def test_agents_sort_keeps_results_by_gid():
    # physical re-ordering by cell: same agents (by gid), same field, as the unsorted oracle run
    cfg = make_cfg(256, 512, 2, 9, D=0.1)
    cfg.fsm = 1
    rng = np.random.default_rng(31)
    n = 20000
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.05)
    st = rng.integers(1, 3, n)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    g = PheroField(cfg)
    sw = Swarm(x, y, th, st, capacity=n + 100)  # trailing empty slots must stay last
    for rounds in range(4):
        g.sort_agents(sw)
        got = sw.to_numpy()
        alive = (got["meta"] & 3) != 0
        assert alive[:n].all() and not alive[n:].any()
        r, q, _ = dense.cells_of(cfg, got["x"][:n], got["y"][:n])
        key = r * cfg.W + q
        assert (np.diff(key) >= 0).all(), "agents not ordered by cell"
        o.tick(a_np, 25)
        g.step(sw, 25)
        got = sw.to_numpy()
        order = np.argsort(got["gid"][:n], kind="stable")
        assert (got["gid"][:n][order] == a_np.gid).all()
        for f in ("x", "y", "theta", "wl", "wr"):
            assert_bitexact(got[f][:n][order], getattr(a_np, f), f)
        assert (got["meta"][:n][order] == a_np.meta).all()
        assert_bitexact(g.snapshot(), o.field, "field after sorted ticks")
    g.close()


def test_graph_replay_equals_plain_launches_and_oracle():
    # small (launch-bound) problems replay pairs of ticks from a CUDA graph; odd/even tick counts,
    # repeated calls (graph reuse) and an agent sort in between (graph re-capture not needed) must
    # all give the oracle's result
    cfg = make_cfg(512, 1024, 2, 5, D=0.08)
    cfg.fsm = 1
    rng = np.random.default_rng(123)
    n = 3000
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.05)
    st = rng.integers(1, 3, n)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    g = PheroField(cfg)
    g.set_option(abi.OPT_USE_GRAPH, 1)
    sw = Swarm(x, y, th, st)
    launches = []
    for k in (7, 12, 6, 33, 1, 2, 9):
        l0 = g.launch_count()
        o.tick(a_np, k)
        g.step(sw, k)
        launches.append(g.launch_count() - l0)
        _compare_agents(sw, a_np, f"after {k} ticks")
        assert_bitexact(g.snapshot(), o.field, f"field after {k} ticks")
    c = g.counters()
    assert c.ticks == 70 and c.deposits == int(o.counters[8])
    assert list(c.decisions) == [int(v) for v in o.counters[:6]]
    # the same run without graphs gives the same state
    g2 = PheroField(cfg)
    g2.set_option(abi.OPT_USE_GRAPH, 0)
    sw2 = Swarm(x, y, th, st)
    g2.step(sw2, 70)
    assert_bitexact(g2.snapshot(), g.snapshot())
    g.close(); g2.close()


def test_owner_computes_deposit_hot_cells_and_long_drift():
    """pf_agents_sort switches the agent deposit to the sort-free owner-computes kernel. Exercise its
    overflow paths (a key range with > 2048 agents is bisected; one cell with > 2048 agents is summed
    sequentially), two channels, and 150 ticks of drift without re-sorting; compare with the oracle
    and with the global-sort path."""
    cfg = make_cfg(256, 512, 2, 5, D=0.05)
    cfg.fsm = 1
    cfg.nest_x, cfg.nest_y = 1.28, 2.56
    cfg.food_x[0], cfg.food_y[0] = 1.9, 3.5
    cfg.food_radius = 0.3
    rng = np.random.default_rng(2024)
    n = 60000
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.05)
    st = rng.integers(1, 3, n)
    # 5000 agents in ONE cell, 9000 more inside a 6 x 6 cell patch (nest-like hot spot)
    x[:5000] = 1.2845; y[:5000] = 2.5655
    x[5000:14000] = (1.30 + rng.uniform(0, 0.06, 9000)).astype(np.float32)
    y[5000:14000] = (2.60 + rng.uniform(0, 0.06, 9000)).astype(np.float32)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    g = PheroField(cfg)
    g.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)  # dense on purpose: the heuristic would pick the radix sort
    g2 = PheroField(cfg)
    g2.set_option(abi.OPT_NO_OWNER_DEPOSIT, 1)
    sw, sw2 = Swarm(x, y, th, st), Swarm(x, y, th, st)
    g.sort_agents(sw)
    g2.sort_agents(sw2)
    for k in (1, 3, 30, 116):
        o.tick(a_np, k)
        g.step(sw, k)
        g2.step(sw2, k)
        assert_bitexact(g.snapshot(), o.field, f"owner-computes field after {k} more ticks")
        assert_bitexact(g2.snapshot(), o.field, f"global-sort field after {k} more ticks")
    got = sw.to_numpy()
    order = np.argsort(got["gid"], kind="stable")
    for f in ("x", "y", "theta", "wl", "wr"):
        assert_bitexact(got[f][order], getattr(a_np, f), f)
    assert (got["meta"][order] == a_np.meta).all()
    assert g.counters().deposits == int(o.counters[8]) == g2.counters().deposits
    g.close(); g2.close()


def test_owner_computes_deposit_standalone_calls_and_fractional_amount_order():
    # pf_agents_deposit (own key kernel) on sorted arrays, non-integer intensities: the per-cell sum
    # order must be the array order, exactly as in the oracle's sequential sum
    cfg = make_cfg(128, 256, 1, 0)
    cfg.intensity_explore, cfg.intensity_high, cfg.intensity_homing = 0.3, 1.7, 0.9
    rng = np.random.default_rng(5)
    n = 30000
    x, y, th, _ = rand_agents(rng, cfg, n)
    x[:8000] = (0.5 + rng.uniform(0, 0.03, 8000)).astype(np.float32)
    y[:8000] = (0.7 + rng.uniform(0, 0.03, 8000)).astype(np.float32)
    st = rng.integers(1, 3, n)
    g = PheroField(cfg)
    g.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    sw = Swarm(x, y, th, st)
    g.sort_agents(sw)
    srt = sw.to_numpy()
    a_np = dense.AgentsNP(srt["x"], srt["y"], srt["theta"], st[srt["gid"]], gid=srt["gid"])
    a_np.meta[:] = srt["meta"]
    o = dense.DenseOracle(cfg)
    for _ in range(4):
        o.agents_deposit(a_np)
        g.agents_deposit(sw)
    assert (sw.to_numpy()["meta"] == a_np.meta).all()
    assert_bitexact(g.snapshot(), o.field, "fractional amounts, owner-computes vs sequential oracle")
    g.close()


# ----------------------------------------------------------------------------- deposit fused into the stencil
@pytest.mark.parametrize("amounts", [(1.0, 20.0, 5.0), (3.0, 7.0, 32768.0), (0.3, 1.7, 0.9)])
@pytest.mark.parametrize("stencil", [5, 9])
def test_deposit_riding_on_the_stencil_bitexact(stencil, amounts):
    # pf_step hands tick t+1's deposits to tick t's stencil (pf_deposit_tile.cuh): 3 x 2 tiles per
    # channel with ragged last tiles (160 = 64 + 64 + 32 rows, 1536 = 1024 + 512 columns), two
    # channels, a few hundred shared cells per tile. Integer amounts (the reference's 1 / 20 / 5, sup:232-234):
    # per-cell totals are exact in any order and ride on the stencil; fractional amounts (summation order
    # matters) keep the separate deposit kernels with their array-order sums — same bits as the oracle either way
    cfg = make_cfg(160, 1536, 2, stencil)
    cfg.fsm = 1
    cfg.intensity_explore, cfg.intensity_high, cfg.intensity_homing = amounts
    rides = all(float(v).is_integer() for v in amounts)
    cfg.nest_x, cfg.nest_y = 0.8, 7.0
    cfg.food_x[0], cfg.food_y[0] = 1.1, 9.0
    cfg.food_radius = 0.3
    rng = np.random.default_rng(77)
    n = 4100   # the two full 64 x 1024 tiles of a channel get about 550 depositors each (limit 768 cells)
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.02)
    x[:100] = (0.40 + rng.uniform(0, 0.06, 100)).astype(np.float32)   # 100 agents on a 6 x 6 cell patch
    y[:100] = (3.00 + rng.uniform(0, 0.06, 100)).astype(np.float32)
    st = rng.integers(1, 3, n)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    g = PheroField(cfg)
    g.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    g2 = PheroField(cfg)
    g2.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    g2.set_option(abi.OPT_NO_TILE_DEPOSIT, 1)
    sw, sw2 = Swarm(x, y, th, st), Swarm(x, y, th, st)
    g.sort_agents(sw)
    g2.sort_agents(sw2)
    ticks = 0
    for k in (1, 2, 5, 40):
        o.tick(a_np, k)
        g.step(sw, k)
        g2.step(sw2, k)
        ticks += k
        assert_bitexact(g.snapshot(), o.field, f"fused field after {ticks} ticks")
        assert_bitexact(g2.snapshot(), o.field, f"unfused field after {ticks} ticks")
    got = sw.to_numpy()
    order = np.argsort(got["gid"], kind="stable")
    for f in ("x", "y", "theta", "wl", "wr"):
        assert_bitexact(got[f][order], getattr(a_np, f), f)
    assert (got["meta"][order] == a_np.meta).all()
    attempted, fell_back = g.tile_stats()
    assert attempted == ((2 - 1) + (5 - 1) + (40 - 1) if rides else 0) and fell_back == 0   # every tick but the last of a call
    assert g2.tile_stats() == (0, 0)
    g.close(); g2.close()


def test_deposit_riding_on_the_stencil_crowded_tile_falls_back_once_then_goes_dense():
    # 5000 agents inside one tile, 1500 of them on a 4 x 4 patch (the bins of its stage groups overflow):
    # the flag is raised on the device, the stencil leaves its output alone and the next
    # tick's own deposit kernels do the work; the tile is marked dense and from then on its amounts go through
    # a per-cell side buffer that the tile's stencil CTA adds to its output — same bits either way
    cfg = make_cfg(128, 2048, 1, 9)
    rng = np.random.default_rng(78)
    n = 9000
    x, y, th, _ = rand_agents(rng, cfg, n, margin=0.02)
    x[:5000] = (0.10 + rng.uniform(0, 0.40, 5000)).astype(np.float32)
    y[:5000] = (12.0 + rng.uniform(0, 4.00, 5000)).astype(np.float32)
    x[:1500] = (0.30 + rng.uniform(0, 0.04, 1500)).astype(np.float32)
    y[:1500] = (14.0 + rng.uniform(0, 0.04, 1500)).astype(np.float32)
    st = rng.integers(1, 3, n)
    a_np = dense.AgentsNP(x, y, th, st)
    o = dense.DenseOracle(cfg)
    g = PheroField(cfg)
    g.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    sw = Swarm(x, y, th, st)
    g.sort_agents(sw)
    for k in (3, 20):
        o.tick(a_np, k)
        g.step(sw, k)
        assert_bitexact(g.snapshot(), o.field, "field with crowded tiles")
    attempted, fell_back = g.tile_stats()
    assert attempted == 2 + 19 and fell_back == 1   # the first hand-over overflowed; the crowded tile is dense since
    assert g.counters().deposits == int(o.counters[8])
    # a re-sort forgets the marks: one more fall-back, then dense again
    g.sort_agents(sw)
    o.tick(a_np, 6)
    g.step(sw, 6)
    assert_bitexact(g.snapshot(), o.field, "field with crowded tiles after a re-sort")
    assert g.tile_stats() == (2 + 19 + 5, 2)
    g.close()
