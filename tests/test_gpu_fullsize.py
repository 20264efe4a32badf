"""BASELINE config 3 at full size (32768 x 32768 cells x 2 channels, 2^24 agents): properties that do not
need the CPU oracle (it would take minutes per tick at this size).

The deposit riding on the stencil must be bit-identical to the separate deposit kernel here too — 32768
stencil CTAs whose producer warps send reductions after the consumer warps' stores, with a field 70 x
the L2: this is where an ordering bug would show, not in the small arenas of test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from swarm_robotics_pheromones_b200 import abi
from swarm_robotics_pheromones_b200.field import PheroField, Swarm

pytestmark = pytest.mark.gpu


def _config3():
    H = W = 32768
    cfg = abi.default_config(H, W, 2)
    cfg.stencil = 9
    cfg.fsm = 1                      # both channels get deposits (SEARCHING -> 0, HOMING -> 1)
    for c in range(2):
        cfg.evap_keep[c] = 0.9
        cfg.diffusion[c] = 0.1
    cfg.nest_x, cfg.nest_y = 0.5 * H * cfg.cell_size, 0.5 * W * cfg.cell_size
    cfg.food_x[0], cfg.food_y[0] = 0.7 * H * cfg.cell_size, 0.6 * W * cfg.cell_size
    cfg.food_radius = 2.0
    return cfg


def _agents(cfg, n, seed=1234):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = (torch.rand(n, generator=g) * (cfg.H * cfg.cell_size)).float().numpy()
    y = (torch.rand(n, generator=g) * (cfg.W * cfg.cell_size)).float().numpy()
    th = ((torch.rand(n, generator=g) * 2 - 1) * float(np.pi)).float().numpy()
    st = torch.randint(1, 3, (n,), generator=g).numpy()
    return x, y, th, st


@pytest.mark.timeout(600)
def test_config3_fused_deposit_equals_separate_deposit_bitexact():
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2**30:
        pytest.skip("needs about 45 GB of device memory")
    cfg = _config3()
    n = 1 << 24
    x, y, th, st = _agents(cfg, n)
    fused, plain = PheroField(cfg), PheroField(cfg)
    plain.set_option(abi.OPT_NO_TILE_DEPOSIT, 1)
    sa, sb = Swarm(x, y, th, st), Swarm(x, y, th, st)
    fused.sort_agents(sa)
    plain.sort_agents(sb)
    for k in (3, 9):                 # two calls: the last tick of a call is never handed over
        fused.step(sa, k)
        plain.step(sb, k)
    torch.cuda.synchronize()
    attempted, fell_back = fused.tile_stats()
    assert attempted == 2 + 8 and fell_back == 0
    assert plain.tile_stats() == (0, 0)
    fa, fb = fused.view(0), plain.view(0)
    assert torch.equal(fa.view(torch.int32), fb.view(torch.int32)), "fields differ"
    for f in ("x", "y", "theta", "wl", "wr", "meta", "gid"):
        assert torch.equal(getattr(sa, f), getattr(sb, f)), f
    ca, cb = fused.counters(), plain.counters()
    assert ca.deposits == cb.deposits > 10 * n and list(ca.decisions) == list(cb.decisions)
    # deposits (1, 20 or 5 per robot and tick) all landed: the field's mass is the decayed sum, nonzero in both channels
    s = fa.sum(dim=(1, 2), dtype=torch.float64)
    assert float(s[0]) > 0 and float(s[1]) > 0
    fused.close(); plain.close()
