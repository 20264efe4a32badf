"""Host-side pieces of bench.py and of the strip partitioning that need no GPU: the cost-balanced strip bounds
(config 4's nest), the synthetic populations, the clock sampler's schedule, the JSON contract of the reference arm."""
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from swarm_robotics_pheromones_b200 import abi, sharded  # noqa: E402


def test_weighted_bounds_cover_the_arena_and_balance_the_cost():
    H = 16384
    rng = np.random.default_rng(0)
    per_row = np.zeros(H)
    per_row[7400:9000] = rng.integers(1000, 3000, 1600)          # a nest: a few rows hold nearly everybody
    cost = 16384 * 2 * bench.BYTES_PER_CELL_UPDATE + per_row * 260.0
    for world in (2, 4, 8):
        b = sharded.weighted_bounds(cost, world, min_rows=256)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == H
        heights = np.diff(b)
        assert (heights >= 256).all()
        per_strip = np.array([cost[b[k]:b[k + 1]].sum() for k in range(world)])
        assert per_strip.max() / per_strip.mean() < 1.35, (world, per_strip)
        # ... which equal heights would not give
        eq = np.array([cost[k * H // world:(k + 1) * H // world].sum() for k in range(world)])
        if world > 2:
            assert per_strip.max() < eq.max()
    # strip_config takes the bounds as given
    cfg, _ = bench.make_config("c4")
    b = sharded.weighted_bounds(cost, 8, min_rows=256)
    for r in range(8):
        s = sharded.strip_config(cfg, 8, r, device=0, bounds=b)
        assert (s.row0, s.row1) == (b[r], b[r + 1])


def test_config4_population_has_searchers_at_the_food_and_carriers_near_the_nest():
    cfg, n = bench.make_config("c4")
    assert cfg.fsm == 1 and cfg.n_food == 8 and (cfg.H, cfg.W, cfg.C) == (16384, 16384, 2)
    n = 1 << 16                                                  # same recipe, fewer robots
    x, y, th, st = bench.init_agents_torch(cfg, n, "cpu")
    k = n // 16
    assert (st[k:2 * k] == abi.STATE_HOMING).all() and (st[:k] == abi.STATE_SEARCHING).all()
    assert (st[2 * k:] == abi.STATE_SEARCHING).all()
    d_nest = np.hypot(x[k:2 * k] - cfg.nest_x, y[k:2 * k] - cfg.nest_y)
    assert d_nest.min() >= 0.039 and d_nest.max() <= 0.701       # inside the home radius (sup:489), outside the stop ring
    fx = np.array([cfg.food_x[i] for i in range(8)])
    fy = np.array([cfg.food_y[i] for i in range(8)])
    d_food = np.min(np.hypot(x[:k, None] - fx[None], y[:k, None] - fy[None]), axis=1)
    assert d_food.max() <= 2.001
    # the rest: nest +- 5 % of the arena width
    assert np.abs(x[2 * k:] - cfg.nest_x).max() <= 0.05 * cfg.H * cfg.cell_size + 1e-3
    rows = sharded.agent_rows(cfg, x)
    assert rows.min() >= 0 and rows.max() < cfg.H
    # configs 1-3: uniform, all searching
    cfg3, _ = bench.make_config("c3")
    x3, y3, _, st3 = bench.init_agents_torch(cfg3, 4096, "cpu")
    assert (st3 == abi.STATE_SEARCHING).all() and 0 <= x3.min() and x3.max() <= cfg3.H * cfg3.cell_size


class _FakeNvml:
    NVML_CLOCK_SM = 0

    def __init__(self):
        self.calls = []

    def nvmlDeviceGetClockInfo(self, h, kind):
        self.calls.append(time.perf_counter())
        return 1965

    def nvmlDeviceGetCurrentClocksThrottleReasons(self, h):
        return 0x4                                               # sw_power_cap: kept and noted


def _sampler(delay):
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    bench.ClockSampler.__init__(s, 0, first_delay_s=delay)       # (NVML may or may not exist here)
    s.nv, s.h, s.max_mhz = _FakeNvml(), None, 1965
    return s


def test_clock_sampler_first_sample_is_delayed_but_never_missing():
    # the first NVML query comes `first_delay_s` into the timed region (eight ranks querying at the region's first
    # instruction stall each other's launches: profiles/r02_notes.md) ...
    s = _sampler(0.05)
    t0 = time.perf_counter()
    s.start()
    time.sleep(0.2)
    out = s.stop()
    assert out["samples"] >= 2 and out["sm_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert s.nv.calls[0] - t0 >= 0.045
    # ... and a region shorter than the delay still gets its one sample
    s = _sampler(0.5)
    s.start()
    time.sleep(0.01)
    out = s.stop()
    assert out["samples"] == 1


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                        "--config", "c1"], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("c1:")
