"""ReferenceCompat: the reference's call shapes over the GPU field, checked against the real
reference's golden trace (decisions + wheel speeds on the shipped trajectory) and the oracle."""
import json

import numpy as np
import pytest
import torch

from oracle import dense
from swarm_robotics_pheromones_b200 import abi
from swarm_robotics_pheromones_b200.compat import ReferenceCompat

pytestmark = pytest.mark.gpu


def c0_config(ps=60.0):
    """testing.wbt geometry: 1 m x 1 m arena centred at 0, 5 mm cells; reference-literal DRUL
    offsets: front = +y, back = -y, left/right = y -+ 1 METRE (sup:399-400), i.e. off-arena."""
    cfg = abi.default_config(200, 200, 1, cell_size=0.005, ps=(ps,) * 8)
    cfg.origin_x, cfg.origin_y = -0.5, -0.5
    cfg.sense_mode = abi.SENSE_WORLD
    drow = (0, 0, 0, 0, 0)
    dcol = (0, 1, -1, -200, 200)
    for d in range(5):
        cfg.sense_drow[d], cfg.sense_dcol[d] = drow[d], dcol[d]
    cfg.food_x[0], cfg.food_y[0] = 0.36, -0.15      # testing.wbt:27
    return cfg


def test_replay_heatmap_matches_reference_decisions_and_speeds(golden_dir):
    g = json.load(open(golden_dir / "sup_replay_heatmap.json"))
    cfg = c0_config(g["ps"])
    cfg.nest_x, cfg.nest_y = g["start"][0], g["start"][1]
    rc = ReferenceCompat(cfg, ["e2"], [g["start"]], start_delay=5.0, time0=g["t0"])
    n_dep = 0
    for rec, pos in zip(g["ticks"], g["positions"]):
        out = rc.startPheromone(["e2"], {"e2": pos})
        wl, wr = out["e2"]
        # reference: always "no change needed" on this trajectory -> STRAIGHT, speeds 2.38778125
        assert rec["direction"] in ("current", "front", "back")
        dec = int((rc.swarm.meta[0].item() & abi.META_DEC_MASK) >> abi.META_DEC_SHIFT)
        assert dec == abi.DEC_STRAIGHT
        assert abs(wl - rec["speeds"][0]) < 1e-6 and abs(wr - rec["speeds"][1]) < 1e-6
        n_dep += rec["amount"] is not None
    c = rc.field.counters()
    # deposit / no-deposit mask (has_moved on the recorded trajectory) and the every-3rd rule
    assert c.deposits == n_dep == g["ticks"][-1]["n_global"]
    count = int(rc.swarm.meta[0].item()) >> abi.META_COUNT_SHIFT
    assert count == g["ticks"][-1]["n_global"]
    rc.close()


def test_replay_homing_matches_reference_decisions(golden_dir):
    g = json.load(open(golden_dir / "sup_replay_homing.json"))
    cfg = abi.default_config(600, 600, 1, cell_size=0.005, ps=(g["ps"],) * 8)
    cfg.origin_x, cfg.origin_y = -1.5, -1.5
    cfg.nest_x, cfg.nest_y = g["start"][0], g["start"][1]
    rc = ReferenceCompat(cfg, ["e2"], [g["positions"][0]], states=abi.STATE_HOMING, start_delay=0.0,
                         time0=g["t0"])
    from oracle import trail
    o = trail.TrailOracle(g["start"], g["food"], ps=(g["ps"],) * 8)
    o.state = trail.HOMING
    names = {"STRAIGHT": abi.DEC_STRAIGHT, "LEFT": abi.DEC_LEFT, "RIGHT": abi.DEC_RIGHT,
             "KEEP": abi.DEC_KEEP, "STOP": abi.DEC_STOP}
    mism = 0
    for rec, pos in zip(g["ticks"], g["positions"]):
        want = o.tick(rec["t"], pos)  # == the real reference (tests/test_oracle_trail.py)
        out = rc.homing(["e2"], {"e2": pos})
        dec = int((rc.swarm.meta[0].item() & abi.META_DEC_MASK) >> abi.META_DEC_SHIFT)
        # fp32 vs fp64 can only disagree where round(turn, 2) or round(d, 2) sits on a boundary
        if dec != names[want["decision"]]:
            mism += 1
            continue
        assert abs(out["e2"][0] - want["speeds"][0]) < 2e-5 * max(1.0, abs(want["speeds"][0]))
        assert abs(out["e2"][1] - want["speeds"][1]) < 2e-5 * max(1.0, abs(want["speeds"][1]))
    assert mism == 0, f"{mism} decision mismatches vs the reference"   # none on the shipped trace
    rc.close()


def test_five_methods_shapes_and_errors():
    cfg = c0_config()
    rc = ReferenceCompat(cfg, ["e1", "e2"], [[0.1, 0.1, 0.0], [-0.2, 0.3, 0.0]], start_delay=0.0)
    rc.deposit_pheromone("e1", [0.1, 0.1, 0.0])
    s = rc.sense_pheromone("e1", [0.1, 0.1, 0.0])
    assert list(s.keys()) == ["current", "front", "back", "left", "right"]  # sup:390 order
    assert s["current"][0][1][0] == 1.0 and s["left"] == [] and s["right"] == []
    wl, wr = rc.adjust_robot_behavior(s, "e1")
    assert abs(wl - 2.38778125) < 1e-6  # v + Braitenberg at ps_j = 60
    # evaporate_pheromone: I * exp(-0.1 t), scalar and array
    v = rc.evaporate_pheromone("e1", 2.048, 20.0)
    assert abs(v - 20.0 * np.exp(-0.1 * 2.048)) < 1e-12
    arr = rc.evaporate_pheromone("e1", np.array([2.0, 3.0]), np.array([1.0, 5.0]), 0.1)
    assert np.allclose(arr, [np.exp(-0.2), 5 * np.exp(-0.3)], rtol=1e-14)
    with pytest.raises(KeyError):  # the reference's dict lookup raises KeyError for unknown robots
        rc.deposit_pheromone("e9", [0, 0, 0])
    # not moved (< 1.5 mm): nothing deposited (sup:321-322)
    before = rc.field.total(0)
    rc.deposit_pheromone("e1", [0.1005, 0.1, 0.0])
    assert rc.field.total(0) == before
    rc.deposit_pheromone("e1", [0.11, 0.1, 0.0])
    assert rc.field.total(0) == before + 1.0
    rc.close()


def test_tick_host_equals_oracle_without_move():
    cfg = abi.default_config(256, 256, 1)
    cfg.stencil = 5
    cfg.diffusion[0] = 0.1
    rng = np.random.default_rng(0)
    n = 5000
    x = rng.uniform(0.1, 2.4, n).astype(np.float32)
    y = rng.uniform(0.1, 2.4, n).astype(np.float32)
    th = rng.uniform(-3, 3, n).astype(np.float32)
    rc = ReferenceCompat(cfg, [str(i) for i in range(n)], np.stack([x, y], 1), headings=th,
                         start_delay=0.0)
    o = dense.DenseOracle(cfg)
    a = dense.AgentsNP(x, y, th)
    pose = torch.empty((3, n)).pin_memory()
    for t in range(20):
        x = (x + rng.uniform(-0.004, 0.004, n)).astype(np.float32)
        pose[0].copy_(torch.from_numpy(x)); pose[1].copy_(torch.from_numpy(y)); pose[2].copy_(torch.from_numpy(th))
        wl, wr, dec = rc.tick_host(pose)
        # oracle: observe (has_moved), deposit, decide, field
        if t == 0:
            moved = np.ones(n, bool)
        else:
            dx = x - a.x
            dy = y - a.y
            moved = np.sqrt(dx * dx + dy * dy, dtype=np.float32) >= np.float32(0.0015)
        a.x[:] = x
        a.meta = (a.meta & ~np.uint32(abi.META_MOVED)) | np.where(moved, abi.META_MOVED, 0).astype(np.uint32) | np.uint32(abi.META_HASPOS)
        o.agents_deposit(a)
        d = o.agents_step(a, abi.STEP_SENSE)
        o.step_field(1)
        assert (dec.numpy() == d).all()
        assert (wl.numpy().view(np.uint32) == a.wl.view(np.uint32)).all()
    assert (rc.field.snapshot().view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    rc.close()


def test_tick_host_pipeline_over_robot_chunks_equals_oracle():
    # 70 000 robots: tick_host runs as a pipeline over four robot sub-ranges (x / y of every chunk first, observe per
    # chunk, one deposit beside the headings' H2D, decide + D2H per chunk); heading-relative sensing on a random
    # two-channel field, so a heading that arrived late or in the wrong chunk would change decisions
    cfg = abi.default_config(512, 512, 2)
    cfg.stencil = 9
    for c in range(2):
        cfg.diffusion[c] = 0.1
    rng = np.random.default_rng(3)
    n = 70000
    assert n >= (1 << 16)
    x = rng.uniform(0.1, 5.0, n).astype(np.float32)
    y = rng.uniform(0.1, 5.0, n).astype(np.float32)
    th = rng.uniform(-3, 3, n).astype(np.float32)
    f0 = rng.random((2, 512, 512), dtype=np.float32)
    rc = ReferenceCompat(cfg, [str(i) for i in range(n)], np.stack([x, y], 1), headings=th, start_delay=0.0)
    assert len(rc._chunk_views()) == 4
    rc.field.upload(f0)
    o = dense.DenseOracle(cfg)
    o.set_field(f0)
    a = dense.AgentsNP(x, y, th)
    pose = torch.empty((3, n)).pin_memory()
    seen = np.zeros(6, np.int64)
    for t in range(6):
        x = (x + rng.uniform(-0.004, 0.004, n)).astype(np.float32)
        y = (y + rng.uniform(-0.004, 0.004, n)).astype(np.float32)
        th = (th + rng.uniform(-0.3, 0.3, n)).astype(np.float32)
        pose[0].copy_(torch.from_numpy(x)); pose[1].copy_(torch.from_numpy(y)); pose[2].copy_(torch.from_numpy(th))
        wl, wr, dec = rc.tick_host(pose)
        if t == 0:
            moved = np.ones(n, bool)
        else:
            dx, dy = x - a.x, y - a.y
            moved = np.sqrt(dx * dx + dy * dy, dtype=np.float32) >= np.float32(0.0015)
        a.x[:], a.y[:], a.theta[:] = x, y, th
        a.meta = (a.meta & ~np.uint32(abi.META_MOVED)) | np.where(moved, abi.META_MOVED, 0).astype(np.uint32) | np.uint32(abi.META_HASPOS)
        o.agents_deposit(a)
        d = o.agents_step(a, abi.STEP_SENSE)
        o.step_field(1)
        assert (dec.numpy() == d).all(), f"tick {t}: {(dec.numpy() != d).sum()} decisions differ"
        assert (wl.numpy().view(np.uint32) == a.wl.view(np.uint32)).all()
        assert (wr.numpy().view(np.uint32) == a.wr.view(np.uint32)).all()
        seen += np.bincount(d, minlength=6)
    assert seen[abi.DEC_LEFT] > 0 and seen[abi.DEC_RIGHT] > 0 and seen[abi.DEC_STRAIGHT] > 0
    got = rc.swarm.to_numpy()
    assert (got["theta"].view(np.uint32) == a.theta.view(np.uint32)).all()
    torch.cuda.synchronize()
    assert (rc.field.snapshot().view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    rc.close()


def test_first_deposit_after_the_start_delay_is_unconditional_for_a_stationary_robot():
    # the reference stores poses only inside deposit_pheromone (sup:279-289, called from t >= 5 s on), so
    # the first deposit has no previous pose and is unconditional (old_pos is None -> True, sup:261-262) — even for a
    # robot that has not moved since t = 0. The second tick then sees no movement: "nothing deposited" (sup:321-322)
    cfg = c0_config()
    rc = ReferenceCompat(cfg, ["e1", "e2"], [[0.1, 0.1, 0.0], [-0.2, 0.3, 0.0]], start_delay=5.0, time0=4.8)
    pose = torch.empty((3, 2)).pin_memory()
    pose[0] = torch.tensor([0.1, -0.2]); pose[1] = torch.tensor([0.1, 0.3]); pose[2] = 0.0
    deposits = []
    for _ in range(6):                      # t = 4.8, 4.864, 4.928, 4.992 (random walk), 5.056, 5.12 (pheromone ticks)
        rc.tick_host(pose)
        torch.cuda.synchronize()
        deposits.append(int(rc.field.counters().deposits))
    assert deposits == [0, 0, 0, 0, 2, 2], deposits
    rc.close()
