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
