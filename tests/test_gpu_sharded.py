"""k row strips (one process, one GPU, LocalDriver) == the unsharded GPU run == the oracle."""
import numpy as np
import pytest
import torch

from oracle import dense
from swarm_robotics_pheromones_b200 import abi, sharded
from swarm_robotics_pheromones_b200.field import PheroField, Swarm
from tests.oracle_strip import gather_agents_sorted
from tests.test_sharded_cpu import _scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("peer", [False, True])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("variant", [1, 2])
def test_strips_equal_unsharded_bitexact(world, variant, peer):
    cfg, x, y, th, st = _scenario(seed=11, n=6000, H=128, W=256, C=2)
    nticks = 150
    o = dense.DenseOracle(cfg)
    a = dense.AgentsNP(x, y, th, st)
    o.tick(a, nticks)
    parts = sharded.partition_agents(cfg, world, x, y, th, st)
    workers = []
    for r in range(world):
        p = parts[r]
        strip = sharded.CudaStrip(sharded.strip_config(cfg, world, r), p["x"], p["y"], p["theta"],
                                  p["state"], p["gid"], capacity=6000, migrate_cap=512)
        strip.field.set_stencil_variant(variant, 8)
        strip.field.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)  # small dense test arena
        workers.append(sharded.StripWorker(strip, r, world))
    drv = sharded.LocalDriver(workers)
    if peer:  # halo fused into the deposit / stencil kernels (peer stores into the neighbours' ghosts)
        drv.attach_peers()
    for t in range(nticks):
        if t % 16 == 0:  # re-order by cell: switches the strips' deposit to the owner-computes kernel,
            for w in workers:  # arrivals then live in the tail chunks until the next sort
                w.b.sort_agents()
        drv.tick(True)
    torch.cuda.synchronize()
    field = np.concatenate([w.b.field.snapshot() for w in workers], axis=1)
    assert (field.view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    got = gather_agents_sorted([w.b.swarm.to_numpy() for w in workers])
    assert len(got["gid"]) == a.n
    for f in dense.AgentsNP.FIELDS:
        assert (got[f].view(np.uint32) == getattr(a, f).view(np.uint32)).all(), f
    migrated = sum(w.b.field.counters().migrated_out for w in workers)
    assert migrated > 0
    assert sum(w.b.field.counters().migrate_dropped for w in workers) == 0
    for w in workers:
        w.b.field.close()


@pytest.mark.parametrize("peer", [False, True, "one-launch", "fold-below", "part-a-1", "no-predeposit", "mig-inline", "edges-side"])
@pytest.mark.parametrize("world", [2, 3])
def test_strips_hand_deposits_to_the_stencil_bitexact(world, peer):
    # strips tall enough to have interior tile rows (64 rows x 1024 columns): the move kernel hands the
    # next deposit of the robots there to the stencil (PF_STEP_HANDOVER), the edge bands and the
    # arrivals keep the separate deposit kernel; both with and without peer-store halos
    cfg, x, y, th, st = _scenario(seed=21, n=14000, H=768, W=512, C=2)
    nticks = 40
    o = dense.DenseOracle(cfg)
    a = dense.AgentsNP(x, y, th, st)
    o.tick(a, nticks)
    parts = sharded.partition_agents(cfg, world, x, y, th, st)
    workers = []
    for r in range(world):
        p = parts[r]
        strip = sharded.CudaStrip(sharded.strip_config(cfg, world, r), p["x"], p["y"], p["theta"],
                                  p["state"], p["gid"], capacity=14000, migrate_cap=1024)
        strip.field.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)  # 4-7 k robots per strip: below the default threshold
        if peer == "one-launch":  # the whole strip in one stencil launch behind the move kernel (PF_OPT_STRIP_ONE_LAUNCH)
            strip.field.set_option(abi.OPT_STRIP_ONE_LAUNCH, 1)
        if peer == "fold-below":  # the handed rows, the band below them and the last row in one launch (value 2)
            strip.field.set_option(abi.OPT_STRIP_ONE_LAUNCH, 2)
        if peer == "no-predeposit":  # the whole deposit at the start of the tick (default: in two parts, see pf_strip_tick)
            strip.field.set_option(abi.OPT_STRIP_NO_PREDEPOSIT, 1)
        if peer == "edges-side":     # two-part deposit with the small stencil launches on the side stream as well
            strip.field.set_option(abi.OPT_STRIP_EDGES_SIDE, 1)
        if peer == "mig-inline":     # migration in line with the stencil launches (and therefore no early deposit either)
            strip.field.set_option(abi.OPT_STRIP_MIG_INLINE, 1)
        if peer == "part-a-1":    # only the first tile row runs before the move kernel; every other row hands over
            strip.field.set_option(abi.OPT_STRIP_PART_A_TILES, 1)
        workers.append(sharded.StripWorker(strip, r, world))
    drv = sharded.LocalDriver(workers)
    if peer:
        drv.attach_peers()
    for t in range(nticks):
        if t % 16 == 0:
            for w in workers:
                w.b.sort_agents()
        # not on the tick before a re-sort, nor on the last one: after a hand-over tick the interior
        # rows already hold the NEXT tick's deposits
        drv.tick(True, hand_over=(t + 1) % 16 != 0 and t != nticks - 1)
    torch.cuda.synchronize()
    field = np.concatenate([w.b.field.snapshot() for w in workers], axis=1)
    assert (field.view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    got = gather_agents_sorted([w.b.swarm.to_numpy() for w in workers])
    assert len(got["gid"]) == a.n
    for f in dense.AgentsNP.FIELDS:
        assert (got[f].view(np.uint32) == getattr(a, f).view(np.uint32)).all(), f
    assert sum(w.b.field.counters().migrated_out for w in workers) > 0
    assert sum(w.b.field.counters().deposits for w in workers) == int(o.counters[8])
    for w in workers:
        attempted, fell_back = w.b.field.tile_stats()
        assert attempted == nticks - 3 and fell_back == 0, (w.rank, attempted, fell_back)
        w.b.field.close()


@pytest.mark.parametrize("peer", [False, True])
def test_strip_hand_over_falls_back_when_an_interior_tile_is_crowded(peer):
    # 5000 robots inside one interior tile of strip 0: its bin overflows, the flag is raised on the device,
    # the stencil leaves its output alone and the NEXT tick's key kernel deposits the handed robots after
    # all (with the amount their already bumped counter says); the tile is marked dense and its later deposits go
    # through a side buffer added by the tile's stencil CTA — same bits as the oracle
    cfg, x, y, th, st = _scenario(seed=23, n=14000, H=768, W=512, C=2)
    rng = np.random.default_rng(5)
    x[:5000] = (0.70 + rng.uniform(0, 0.50, 5000)).astype(np.float32)   # rows 70..120 of strip 0: inside tile row 1
    y[:5000] = (1.00 + rng.uniform(0, 2.00, 5000)).astype(np.float32)
    nticks, world = 12, 2
    o = dense.DenseOracle(cfg)
    a = dense.AgentsNP(x, y, th, st)
    o.tick(a, nticks)
    parts = sharded.partition_agents(cfg, world, x, y, th, st)
    workers = []
    for r in range(world):
        p = parts[r]
        strip = sharded.CudaStrip(sharded.strip_config(cfg, world, r), p["x"], p["y"], p["theta"],
                                  p["state"], p["gid"], capacity=16000, migrate_cap=1024)
        strip.field.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
        workers.append(sharded.StripWorker(strip, r, world))
    drv = sharded.LocalDriver(workers)
    if peer:  # pf_strip_tick: the deposit in two parts — the LATE part (handed robots) has real work on the fall-back tick
        drv.attach_peers()
    for w in workers:
        w.b.sort_agents()
    for t in range(nticks):
        drv.tick(True, hand_over=t != nticks - 1)
    torch.cuda.synchronize()
    field = np.concatenate([w.b.field.snapshot() for w in workers], axis=1)
    assert (field.view(np.uint32) == np.ascontiguousarray(o.field).view(np.uint32)).all()
    got = gather_agents_sorted([w.b.swarm.to_numpy() for w in workers])
    for f in dense.AgentsNP.FIELDS:
        assert (got[f].view(np.uint32) == getattr(a, f).view(np.uint32)).all(), f
    assert sum(w.b.field.counters().deposits for w in workers) == int(o.counters[8])
    s0, s1 = workers[0].b.field.tile_stats(), workers[1].b.field.tile_stats()
    assert s0 == (nticks - 1, 1), s0              # crowded strip: the first hand-over fell back, then the tile is dense
    assert s1 == (nticks - 1, 0), s1              # its neighbour: none did
    for w in workers:
        w.b.field.close()


def test_peer_wait_is_bounded_and_reports_a_dead_neighbour():
    # strip 0 ticks alone: its neighbour never signals. The wait gives up after the configured bound instead of
    # wedging the stream, and pf_peer_status / pf_counters report it
    from swarm_robotics_pheromones_b200._ffi import PFError
    cfg, x, y, th, st = _scenario(seed=3, n=2000, H=128, W=256, C=1)
    parts = sharded.partition_agents(cfg, 2, x, y, th, st)
    strips = [sharded.CudaStrip(sharded.strip_config(cfg, 2, r), parts[r]["x"], parts[r]["y"], parts[r]["theta"],
                                parts[r]["state"], parts[r]["gid"], capacity=3000, migrate_cap=256) for r in range(2)]
    strips[0].peer_attach_local(None, strips[1])
    strips[1].peer_attach_local(strips[0], None)
    strips[0].field.set_option(abi.OPT_PEER_TIMEOUT_MS, 50)
    strips[0].field.peer_status()                       # nothing yet
    strips[0].strip_tick(sharded.strip_flags(True, True, False))
    torch.cuda.synchronize()
    with pytest.raises(PFError) as e:
        strips[0].field.peer_status()
    assert e.value.code == abi.PF_ESTATE and "timed out" in str(e.value)
    assert strips[0].field.counters().peer_timeouts >= 1
    for s_ in strips:
        s_.field.close()


def test_snapshot_and_counters_refuse_the_state_after_a_hand_over_tick():
    # after a hand-over tick the interior rows and the deposit counters are one tick ahead; reading
    # them is refused until a tick without PF_STEP_HANDOVER has run
    from swarm_robotics_pheromones_b200._ffi import PFError
    cfg, x, y, th, st = _scenario(seed=21, n=14000, H=768, W=512, C=2)
    parts = sharded.partition_agents(cfg, 2, x, y, th, st)
    workers = []
    for r in range(2):
        p = parts[r]
        strip = sharded.CudaStrip(sharded.strip_config(cfg, 2, r), p["x"], p["y"], p["theta"], p["state"], p["gid"],
                                  capacity=14000, migrate_cap=1024)
        strip.field.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
        workers.append(sharded.StripWorker(strip, r, 2))
    drv = sharded.LocalDriver(workers)
    drv.attach_peers()
    for w in workers:
        w.b.sort_agents()
    drv.tick(True, hand_over=True)
    with pytest.raises(PFError):
        workers[0].b.field.snapshot()
    with pytest.raises(PFError):
        workers[0].b.field.counters()
    with pytest.raises(PFError):
        workers[0].b.sort_agents()
    drv.tick(True)                                      # default: no hand-over
    workers[0].b.field.snapshot()
    workers[0].b.field.counters()
    for w in workers:
        w.b.field.close()
