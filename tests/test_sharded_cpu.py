"""Host-side sharding logic on CPU: strip bounds, partitioning, and the full tick protocol
(halo exchange + agent migration) with world_size-2 gloo processes and the oracle as compute
backend; k strips must equal the unsharded oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense
from swarm_robotics_pheromones_b200 import abi, sharded
from tests.oracle_strip import OracleStrip, gather_agents_sorted


def test_strip_bounds_cover_and_balance():
    for H in (1, 7, 64, 1000, 32768):
        for world in (1, 2, 3, 4, 8):
            if world > H:
                continue
            b = [sharded.strip_bounds(H, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == H
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [r1 - r0 for r0, r1 in b]
            assert max(sizes) - min(sizes) <= 1


def test_owner_and_partition():
    cfg = abi.default_config(64, 32, 1)
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 0.64, 1000).astype(np.float32)
    y = rng.uniform(0, 0.32, 1000).astype(np.float32)
    parts = sharded.partition_agents(cfg, 4, x, y, np.zeros(1000, np.float32))
    assert sum(len(p["x"]) for p in parts) == 1000
    for r, p in enumerate(parts):
        r0, r1 = sharded.strip_bounds(64, 4, r)
        rows = sharded.agent_rows(cfg, p["x"])
        assert ((rows >= r0) & (rows < r1)).all()
    # the row rule is the oracle's
    rr, _, _ = dense.cells_of(cfg, x, y)
    assert (rr == sharded.agent_rows(cfg, x)).all()


def _scenario(seed=3, n=3000, H=96, W=64, C=2):
    cfg = abi.default_config(H, W, C)
    cfg.stencil = 9
    for c in range(C):
        cfg.diffusion[c] = 0.1
        cfg.evap_keep[c] = 0.95
    cfg.fsm = 1
    cfg.nest_x, cfg.nest_y = 0.48, 0.32
    cfg.food_x[0], cfg.food_y[0] = 0.8, 0.5
    cfg.food_radius = 0.08
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.02, H * 0.01 - 0.02, n).astype(np.float32)
    y = rng.uniform(0.02, W * 0.01 - 0.02, n).astype(np.float32)
    th = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    st = rng.integers(1, 3, n)
    return cfg, x, y, th, st


def _reference_run(cfg, x, y, th, st, nticks):
    o = dense.DenseOracle(cfg)
    a = dense.AgentsNP(x, y, th, st)
    o.tick(a, nticks)
    return o, a


def _make_worker(cfg, world, rank, parts, cap_factor=2.0, mcap=512):
    scfg = sharded.strip_config(cfg, world, rank)
    p = parts[rank]
    cap = int(len(parts[0]["x"]) * cap_factor) + 64
    b = OracleStrip(scfg, p["x"], p["y"], p["theta"], p["state"], p["gid"], cap, mcap)
    return sharded.StripWorker(b, rank, world)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_local_driver_equals_unsharded(world):
    cfg, x, y, th, st = _scenario()
    nticks = 120
    o, a = _reference_run(cfg, x, y, th, st, nticks)
    parts = sharded.partition_agents(cfg, world, x, y, th, st)
    workers = [_make_worker(cfg, world, r, parts) for r in range(world)]
    drv = sharded.LocalDriver(workers)
    for _ in range(nticks):
        drv.tick(True)
    field = np.concatenate([w.b.o.field for w in workers], axis=1)
    assert (field.view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    got = gather_agents_sorted([{f: getattr(w.b.a, f) for f in dense.AgentsNP.FIELDS} for w in workers])
    assert len(got["gid"]) == a.n
    for f in dense.AgentsNP.FIELDS:
        want = getattr(a, f)
        assert (got[f].view(np.uint32) == want.view(np.uint32)).all(), f
    # agents did migrate, and every agent sits in its owner strip
    for w in workers:
        alive = (w.b.a.meta & 3) != 0
        rows = sharded.agent_rows(cfg, w.b.a.x[alive])
        assert ((rows >= w.b.cfg.row0) & (rows < w.b.cfg.row1)).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_rank(rank, world, port, nticks, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, x, y, th, st = _scenario()
    parts = sharded.partition_agents(cfg, world, x, y, th, st)
    w = _make_worker(cfg, world, rank, parts)
    drv = sharded.DistDriver(w)
    for _ in range(nticks):
        drv.tick(True)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), field=w.b.o.field,
             **{f: getattr(w.b.a, f) for f in dense.AgentsNP.FIELDS})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_equals_unsharded(tmp_path):
    world, nticks = 2, 80
    cfg, x, y, th, st = _scenario()
    o, a = _reference_run(cfg, x, y, th, st, nticks)
    mp.spawn(_gloo_rank, args=(world, _free_port(), nticks, str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    field = np.concatenate([r["field"] for r in ranks], axis=1)
    assert (field.view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    got = gather_agents_sorted([{f: r[f] for f in dense.AgentsNP.FIELDS} for r in ranks])
    for f in dense.AgentsNP.FIELDS:
        assert (got[f].view(np.uint32) == getattr(a, f).view(np.uint32)).all(), f


def test_worker_row_schedule_covers_every_row_exactly_once():
    # the stencil launches of one tick, with and without a hand-over range: every local row exactly once,
    # the hand-over rows as ONE call after the move kernel, nothing but rows [1, r0) before it
    class Rec:
        def __init__(self, rows, dep):
            self.rows, self._dep, self.calls, self.flags = rows, dep, [], None
        def tile_rows(self):
            return self._dep
        def field_rows(self, r0, r1):
            self.calls.append(("rows", r0, r1))
        def agents(self, flags):
            self.flags = flags
            self.calls.append(("agents",))
        def swap(self):
            self.calls.append(("swap",))

    for rows, dep in [(384, (64, 320)), (384, (64, 384)), (256, (0, 0)), (4096, (512, 4032)), (1, (0, 0)), (2, (0, 0))]:
        b = Rec(rows, dep)
        w = sharded.StripWorker(b, 1, 3)
        w.begin_tick(True)
        w.field_interior_a()
        w.agents(True)
        w.field_rest()
        seen = np.zeros(rows, dtype=int)
        before_agents = True
        for c in b.calls:
            if c[0] == "agents":
                before_agents = False
            elif c[0] == "rows":
                seen[c[1]:c[2]] += 1
                if before_agents:
                    assert c[1] == 1 and (dep[1] <= dep[0] or c[2] == dep[0])
        assert (seen == 1).all(), (rows, dep, b.calls)
        if dep[1] > dep[0]:
            assert ("rows", dep[0], dep[1]) in b.calls and b.flags & abi.STEP_HANDOVER
        else:
            assert not (b.flags & abi.STEP_HANDOVER)
        assert b.calls[-1] == ("swap",)
        # no hand-over asked for: the old half / half split
        b2 = Rec(rows, dep)
        w2 = sharded.StripWorker(b2, 1, 3)
        w2.begin_tick(False)
        w2.field_interior_a(); w2.agents(True); w2.field_rest()
        seen = np.zeros(rows, dtype=int)
        for c in b2.calls:
            if c[0] == "rows":
                seen[c[1]:c[2]] += 1
        assert (seen == 1).all() and not (b2.flags & abi.STEP_HANDOVER)


def _make_ho_worker(cfg, world, rank, parts):
    scfg = sharded.strip_config(cfg, world, rank)
    p = parts[rank]
    b = OracleStrip(scfg, p["x"], p["y"], p["theta"], p["state"], p["gid"], len(p["x"]) * 2 + 64, 512,
                    hand_over=True)
    return sharded.StripWorker(b, rank, world)


def _ho_scenario():
    return _scenario(seed=9, n=2500, H=640, W=64, C=2)   # strips of 320 / 213 rows: interior 64-row bands exist


@pytest.mark.parametrize("world", [2, 3])
def test_local_driver_with_hand_over_equals_unsharded(world):
    # the hand-over protocol (pf_tile_rows / PF_STEP_HANDOVER) driven by the same worker code as on the GPU,
    # with the oracle backend emulating what the library does
    cfg, x, y, th, st = _ho_scenario()
    nticks = 60
    o, a = _reference_run(cfg, x, y, th, st, nticks)
    parts = sharded.partition_agents(cfg, world, x, y, th, st)
    workers = [_make_ho_worker(cfg, world, r, parts) for r in range(world)]
    drv = sharded.LocalDriver(workers)
    for t in range(nticks):
        drv.tick(True, hand_over=t != nticks - 1 and t % 7 != 3)   # some ticks without, never the last one
    field = np.concatenate([w.b.o.field for w in workers], axis=1)
    assert (field.view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    got = gather_agents_sorted([{f: getattr(w.b.a, f) for f in dense.AgentsNP.FIELDS} for w in workers])
    for f in dense.AgentsNP.FIELDS:
        assert (got[f].view(np.uint32) == getattr(a, f).view(np.uint32)).all(), f
    assert all(w.b.handed_ticks == sum(1 for t in range(nticks) if t != nticks - 1 and t % 7 != 3) for w in workers)
    assert sum(int(w.b.o.counters[8]) for w in workers) == int(o.counters[8])


def _gloo_rank_ho(rank, world, port, nticks, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, x, y, th, st = _ho_scenario()
    parts = sharded.partition_agents(cfg, world, x, y, th, st)
    w = _make_ho_worker(cfg, world, rank, parts)
    drv = sharded.DistDriver(w)
    for t in range(nticks):
        drv.tick(True, hand_over=t != nticks - 1)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), field=w.b.o.field, handed=w.b.handed_ticks,
             **{f: getattr(w.b.a, f) for f in dense.AgentsNP.FIELDS})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_with_hand_over_equals_unsharded(tmp_path):
    world, nticks = 2, 40
    cfg, x, y, th, st = _ho_scenario()
    o, a = _reference_run(cfg, x, y, th, st, nticks)
    mp.spawn(_gloo_rank_ho, args=(world, _free_port(), nticks, str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    field = np.concatenate([r["field"] for r in ranks], axis=1)
    assert (field.view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    got = gather_agents_sorted([{f: r[f] for f in dense.AgentsNP.FIELDS} for r in ranks])
    for f in dense.AgentsNP.FIELDS:
        assert (got[f].view(np.uint32) == getattr(a, f).view(np.uint32)).all(), f
    assert all(int(r["handed"]) == nticks - 1 for r in ranks)
