#!/usr/bin/env python
"""bench.py — whole-tick throughput of the pheromone hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU arm: the oracle port on host cores)

One "step" = one supervisor tick over the whole arena and every agent, in the reference order
deposit -> sense/decide -> evaporate+diffuse -> move. Workload (BASELINE.json configs[3], the one
the metric and the north_star target are quoted on; it fits one GPU): 32768 x 32768 cells, two
channels, 2^24 agents, uniform, fp32. At N > 1 the same global arena is split into row strips
(strong scaling), one halo row per channel per internal boundary per tick plus agent migration.

Prints ONE JSON line. `value` = grid cell-updates/s with everything resident in HBM (agent-steps/s
is reported next to it); `e2e` = the same tick driven through the reference-facing call with the
poses coming from pinned HOST memory and the wheel speeds/decisions going back to the host every
step; `roofline` = the evaporate+diffuse kernel's algorithmic bytes over its device time (CUDA
events recorded on the launching stream inside the timed region) against the measured HBM copy
rate; `cpu_baseline` = the C oracle port timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from swarm_robotics_pheromones_b200 import abi  # noqa: E402

CONFIGS = {
    # name: (H, W, C, agents, stencil)
    "c1": (1024, 1024, 1, 4096, 5),
    "c2": (8192, 8192, 1, 1 << 20, 9),
    "c3": (32768, 32768, 2, 1 << 24, 9),
    "c4": (16384, 16384, 2, 1 << 22, 9),
}
BYTES_PER_CELL_UPDATE = 8          # SURVEY §8d / BASELINE.md §5
BYTES_PER_AGENT_STEP = 56


def make_config(name: str, H=None, W=None, C=None, agents=None, stencil=None):
    h, w, c, n, st = CONFIGS[name]
    H, W, C, n, st = H or h, W or w, C or c, agents or n, stencil if stencil is not None else st
    cfg = abi.default_config(H, W, C)
    cfg.stencil = st
    for k in range(C):
        cfg.evap_keep[k] = 1.0 - 0.1
        cfg.diffusion[k] = 0.1
    if name == "c4":  # nest disc + 8 food sources on a ring at 0.35 W
        cfg.fsm = 1
        cx, cy = 0.5 * H * cfg.cell_size, 0.5 * W * cfg.cell_size
        cfg.nest_x, cfg.nest_y = cx, cy
        cfg.n_food = 8
        for k in range(8):
            cfg.food_x[k] = cx + 0.35 * W * cfg.cell_size * np.cos(2 * np.pi * k / 8)
            cfg.food_y[k] = cy + 0.35 * W * cfg.cell_size * np.sin(2 * np.pi * k / 8)
        cfg.food_radius = 0.16
    return cfg, n


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def stencil_traffic_from_profile():
    """dram bytes per launch of the stencil kernel from the committed ncu summary, if any"""
    p = ROOT / "profiles" / "stencil_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index: int, first_delay_s: float = 0.004):
        self.index = index
        self.first_delay_s = float(os.environ.get("PF_BENCH_SAMPLER_DELAY_S", first_delay_s))
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        # NVML queries take the driver's global lock (every rank's sampler at the same instant, right behind the
        # barrier, in front of the first timed tick): the first sample is taken a few milliseconds into the timed
        # region instead of at its first instruction
        self._stop.wait(self.first_delay_s)
        while True:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            if self._stop.is_set():
                break
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_throughput(cfg_full, n_agents_full, budget_s=12.0, threads=None, stencil=None):
    """Time the C oracle on host cores on a bounded sample: each thread owns an independent arena of
    `rows_s` x W x C cells with the same agent density, and runs whole ticks."""
    from oracle import dense
    dense.lib()
    T = threads or max(1, min(os.cpu_count() or 1, 64))
    W = cfg_full.W
    rows_s = max(64, min(cfg_full.H, (1 << 24) // max(1, W * cfg_full.C)))  # ~16 M cells/thread
    n_s = max(1, int(n_agents_full * (rows_s / cfg_full.H)))
    cfg = abi.default_config(rows_s, W, cfg_full.C)
    cfg.stencil = cfg_full.stencil
    for k in range(cfg_full.C):
        cfg.evap_keep[k] = cfg_full.evap_keep[k]
        cfg.diffusion[k] = cfg_full.diffusion[k]
    results = [0] * T
    worlds = []
    for t in range(T):
        rng = np.random.default_rng(1234 + t)
        a = dense.AgentsNP(rng.uniform(0, rows_s * cfg.cell_size, n_s),
                           rng.uniform(0, W * cfg.cell_size, n_s), rng.uniform(-np.pi, np.pi, n_s))
        worlds.append((dense.DenseOracle(cfg), a))
    worlds[0][0].tick(worlds[0][1], 1)  # warm-up / page-in
    stop_at = time.perf_counter() + budget_s

    def work(t):
        o, a = worlds[t]
        k = 0
        while True:
            o.tick(a, 1)  # ctypes releases the GIL
            k += 1
            if time.perf_counter() >= stop_at:
                break
        results[t] = k

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    ticks = sum(results)
    cells = ticks * rows_s * W * cfg.C
    return {
        "value": cells / dt, "unit": "cell-updates/s", "agent_steps_per_s": ticks * n_s / dt,
        "cores": T, "kind": "port",
        "sample": f"{T} threads x independent {rows_s}x{W}x{cfg.C} arenas with {n_s} agents each "
                  f"(same density and stencil as the workload), {ticks} whole ticks in {dt:.1f} s, "
                  f"oracle/pf_oracle.c (gcc -O2, scalar)",
    }, dt / max(1, ticks) * T


def cpu_reference_legs(budget_s=9.0):
    """the reference's own Python arithmetic, timed single-threaded on this box's host
    (the reference has no parallelism). The reference checkout does not travel to the GPU box, so each leg is the
    few lines it consists of, stated here with their file:line, on the reference's own sizes:
      algo_literal   Phase-3/pheromone_algo.py:351-356 — deposit into one cell, then `grid[i][j] *= (1 - rate)` over
                     every cell of a list of lists, at 10 x 10 (as shipped) and 1024 x 1024
      graph_loop     Phase-3/graph.py:17-24 — `heatmap[..] += 20; heatmap = gaussian_filter(heatmap, 5)` per
                     trajectory point on the 497 x 495 float64 map (scipy), points of Phase-3/heatmap.txt
      sup_trail      Phase-3/pheromone_algo_supervisor.py:565-607 — one supervisor tick (deposit, sense, adjust,
                     age-window evaporation) per recorded pose of the shipped trajectory, through oracle/trail.py
                     (the f64 restatement that equals the real reference's traces exactly; kind "port")
    """
    out = {"cores": 1, "note": "single thread; reference figures measured on the real files in the build container: "
                               "SURVEY.md §6 (8-21 M cell-updates/s, 850 agent-steps/s)"}
    share = budget_s / 4.0

    def algo_literal(N):
        grid = [[0.0] * N for _ in range(N)]
        intensity, rate = 1.0, 0.1
        t0 = time.perf_counter()
        k = 0
        while True:
            grid[3 % N][4 % N] += intensity               # algo:351
            for i in range(N):                            # algo:354-356
                row = grid[i]
                for j in range(N):
                    row[j] *= (1 - rate)
            k += 1
            if time.perf_counter() - t0 >= share or (N <= 16 and k >= 20000):
                break
        dt = time.perf_counter() - t0
        return {"grid": N, "steps": k, "ms_per_step": dt / k * 1e3, "cell_updates_per_s": N * N * k / dt}

    out["algo_literal_10"] = algo_literal(10)
    out["algo_literal_1024"] = algo_literal(1024)
    try:
        from scipy.ndimage import gaussian_filter
        g = json.loads((ROOT / "tests" / "golden" / "sup_replay_heatmap.json").read_text())
        H, W = 497, 495                                   # background.png, graph.py:7-10
        heat = np.zeros((H, W))
        t0 = time.perf_counter()
        k = 0
        for p in g["positions"]:
            r, c = abs(int(p[1] * H)), abs(int(p[0] * W))  # graph.py:19-22
            heat[r, c] += 20                               # graph.py:23
            heat = gaussian_filter(heat, sigma=5)          # graph.py:24
            k += 1
            if time.perf_counter() - t0 >= share:
                break
        dt = time.perf_counter() - t0
        out["graph_loop_497x495"] = {"points": k, "ms_per_step": dt / k * 1e3, "cell_updates_per_s": H * W * k / dt,
                                     "mass": float(heat.sum()), "mass_expected": 20.0 * k}
        from oracle import trail
        o = trail.TrailOracle(g["start"], g["food"], ps=(g["ps"],) * 8)
        t0 = time.perf_counter()
        k = 0
        for rec, pos in zip(g["ticks"], g["positions"]):
            o.tick(rec["t"], pos)
            k += 1
            if time.perf_counter() - t0 >= share:
                break
        dt = time.perf_counter() - t0
        out["sup_trail_c0_replay"] = {"ticks": k, "ms_per_tick": dt / k * 1e3, "agent_steps_per_s": k / dt,
                                      "kind": "port (oracle/trail.py, exact vs the reference's golden traces)"}
    except Exception as e:  # a missing fixture must not cost the headline
        out["error"] = repr(e)
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, n = make_config(args.config, args.grid, args.grid, args.channels, args.agents, args.stencil)
    budget = max(5.0, min(60.0, 4.0 * (args.steps + args.warmup)))
    cb, ms = cpu_port_throughput(cfg, n, budget_s=budget)
    line = {
        "impl": "reference", "metric": "grid cell-updates/s (agent-steps/s alongside)",
        "value": cb["value"], "unit": "cell-updates/s", "agent_steps_per_s": cb["agent_steps_per_s"],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_dict(args, cfg, n),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_dict(args, cfg, n):
    return {"workload": f"{args.config}: {cfg.H}x{cfg.W} cells x {cfg.C} channel(s), {n} agents, "
                        f"{cfg.stencil}-point evaporate+diffuse, deposit, sense/decide/move",
            "grid": [cfg.H, cfg.W], "channels": cfg.C, "agents": n, "stencil": cfg.stencil,
            "parallelism": f"row strips x{args.gpus}" if args.gpus > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (field >> 126 MB)" if cfg.H * cfg.W * cfg.C * 4 > (1 << 28)
            else "field fits L2: L2 flushed between timed steps is NOT applied; reported as L2-resident"}


# ------------------------------------------------------------------------------------------ GPU arm
def init_agents_torch(cfg, n, device, seed=1234):
    """(x, y, theta, state) of the synthetic population (torch Philox on the host, fixed seed — used only for the
    INITIAL poses; the reference has no random draws).
    Uniform over the arena, all SEARCHING, for configs 1-3. Config 4 (fsm on): SURVEY §8d's nest cluster (nest +-
    5 % of the arena width, SEARCHING) for 7/8 of the robots; the arena is 164 m wide and a robot covers 3.3 mm per
    tick, so in a bench-length run nobody would get from the nest to a food source or back — the outcome counters
    would stay at zero. The remaining 1/8 is therefore a slice of a run in progress: 1/16 SEARCHING within 2 m of
    the eight food sources (those crossing a food disc grab), 1/16 HOMING carriers 4-70 cm from the nest (inside
    the home radius: they steer by sup:489-500 and deliver in the 5-6 cm window, sup:665-670)."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    xmax, ymax = cfg.H * cfg.cell_size, cfg.W * cfg.cell_size
    x = torch.rand(n, generator=g) * xmax
    y = torch.rand(n, generator=g) * ymax
    th = (torch.rand(n, generator=g) * 2 - 1) * float(np.pi)
    st = torch.full((n,), abi.STATE_SEARCHING, dtype=torch.int64)
    if cfg.fsm:
        x = cfg.nest_x + (torch.rand(n, generator=g) * 2 - 1) * 0.05 * xmax
        y = cfg.nest_y + (torch.rand(n, generator=g) * 2 - 1) * 0.05 * ymax
        k = n // 16
        if k > 0:
            r = 2.0 * torch.sqrt(torch.rand(k, generator=g))
            a = torch.rand(k, generator=g) * 2 * float(np.pi)
            src = torch.randint(0, cfg.n_food, (k,), generator=g)
            fx = torch.tensor([cfg.food_x[i] for i in range(cfg.n_food)])
            fy = torch.tensor([cfg.food_y[i] for i in range(cfg.n_food)])
            x[:k] = fx[src] + r * torch.cos(a)
            y[:k] = fy[src] + r * torch.sin(a)
            r2 = 0.04 + 0.66 * torch.rand(k, generator=g)
            a2 = torch.rand(k, generator=g) * 2 * float(np.pi)
            x[k:2 * k] = cfg.nest_x + r2 * torch.cos(a2)
            y[k:2 * k] = cfg.nest_y + r2 * torch.sin(a2)
            st[k:2 * k] = abi.STATE_HOMING
    return x.float().numpy(), y.float().numpy(), th.float().numpy(), st.numpy()


def measure_other_config(name, steps, sort_every=16):
    """short device-resident run of another BASELINE config (reported next to the headline)"""
    import torch

    from swarm_robotics_pheromones_b200.field import PheroField, Swarm
    cfg, n = make_config(name)
    x, y, th, st = init_agents_torch(cfg, n, "cuda")
    f = PheroField(cfg)
    sw = Swarm(x, y, th, st)
    f.reserve_agents(n)
    cells = cfg.H * cfg.W * cfg.C
    if cells <= (1 << 24):
        sort_every = steps  # the field lives in L2: re-ordering the agents buys nothing
    f.sort_agents(sw)
    f.step(sw, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = 0
    while done < steps:
        m = min(sort_every, steps - done)
        if cells > (1 << 24):
            f.sort_agents(sw)
        f.step(sw, m)
        done += m
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    c = f.counters()
    att, fb = f.tile_stats()
    out = {"fused_deposit": {"ticks_attempted": att, "ticks_fell_back": fb, "dense_tiles": f.tile_dense()[0]},
           "workload": f"{name}: {cfg.H}x{cfg.W}x{cfg.C}, {n} agents, {cfg.stencil}-point",
           "steps": steps, "ms_per_step": ms, "cell_updates_per_s": cells / (ms * 1e-3),
           "agent_steps_per_s": n / (ms * 1e-3),
           "whole_tick_GBs": (cells * 8 + n * 76 + 8 * min(n, cells)) / (ms * 1e-3) / 1e9}
    if cfg.fsm:
        from swarm_robotics_pheromones_b200.metrics import NetEnergy
        out["food_grabbed"], out["food_delivered"] = int(c.food_grabbed), int(c.food_delivered)
        out["net_energy"] = NetEnergy().update(c.food_delivered, c.ticks, n)  # E += F*E_f - E_c (PDF p.19-21)
    f.close()
    return out


def measure_trail_engine(n_robots=1 << 20, ticks=50):
    """sparse timestamped-trail engine (the shipped supervisor's data structure), robots x ticks / s"""
    import torch

    from swarm_robotics_pheromones_b200.trail import TrailSwarm
    names = [str(i) for i in range(n_robots)]
    rng = np.random.default_rng(0)
    start = rng.uniform(-0.5, 0.5, (n_robots, 3))
    start[:, 2] = 0.0
    sw = TrailSwarm(names, start, [0.36, -0.15, 0.05], ps=60.0, start_delay=5.0, record_sense=False)
    pos = torch.from_numpy(start).cuda()
    step = torch.from_numpy(rng.normal(0, 0.003, (n_robots, 3)) * np.array([1, 1, 0])).cuda()
    t = 5.0
    for _ in range(40):  # fill the trails to steady state (about 44 live entries)
        pos += step
        sw.tick_device(t, pos)
        t += 0.064
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ticks):
        pos += step
        sw.tick_device(t, pos)
        t += 0.064
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / ticks
    live = float(sw.out["n_live"].float().mean().item())
    sw.close()
    return {"robots": n_robots, "ms_per_tick": ms, "robot_ticks_per_s": n_robots / (ms * 1e-3),
            "mean_live_entries": live,
            "reference_python_robot_ticks_per_s": 850.0,
            "note": "reference figure: SURVEY §6, headless, 1 core, this container"}


def run_single(args):
    import torch

    from swarm_robotics_pheromones_b200.compat import ReferenceCompat
    from swarm_robotics_pheromones_b200.field import PheroField, Swarm

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(0)
    cfg, n = make_config(args.config, args.grid, args.grid, args.channels, args.agents, args.stencil)
    x, y, th, st0 = init_agents_torch(cfg, n, "cuda")
    field = PheroField(cfg)
    field.set_stencil_variant(args.variant, args.rows_per_cta)
    if args.overlap_agents:
        field.set_option(abi.OPT_OVERLAP_AGENTS, 1)
    if args.force_owner_deposit:
        field.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    swarm = Swarm(x, y, th, st0)
    field.reserve_agents(n)
    field.set_time(0.0, 0.0)
    cells = cfg.H * cfg.W * cfg.C
    if cells <= (1 << 24):
        args.sort_every = 0  # the field lives in L2: re-ordering the agents buys nothing

    def run_ticks(k):
        """k whole ticks; every --sort-every ticks the agent arrays are re-ordered by cell (inside
        the timed region: it is part of the cost of keeping the gathers local)"""
        done = 0
        while done < k:
            if args.sort_every > 0 and state["since_sort"] >= args.sort_every:
                field.sort_agents(swarm)
                state["since_sort"] = 0
                state["sorts"] += 1
            m = k - done
            if args.sort_every > 0:
                m = min(m, args.sort_every - state["since_sort"])
            field.step(swarm, m)
            done += m
            state["since_sort"] += m

    state = {"since_sort": args.sort_every, "sorts": 0}  # first call sorts the initial population

    # ---- device-resident throughput
    sampler = ClockSampler(0)   # (NVML initialised before the warm-up: no idle gap between warm-up and timed region)
    run_ticks(args.warmup)
    torch.cuda.synchronize()
    l0 = field.launch_count()
    # launch-bound (small) problems replay ticks from a CUDA graph, which per-phase event profiling
    # would disable: time them unprofiled and take the phase split from a second, profiled pass
    small = cells <= (1 << 24)
    if not small:
        field.profile(True)
        field.profile_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    torch.cuda.synchronize()
    sorts0 = state["sorts"]
    state["since_sort"] = args.sort_every  # the timed region pays ceil(K / sort_every) sorts
    e0.record()
    run_ticks(args.steps)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    sorts_timed = state["sorts"] - sorts0
    launches = field.launch_count() - l0
    if small:
        field.profile(True)
        field.profile_read(reset=True)
        run_ticks(min(args.steps, 64))
        torch.cuda.synchronize()
    prof = field.profile_read(reset=True)
    field.profile(False)
    tile_att, tile_fb = field.tile_stats()
    tile_dense = field.tile_dense()
    ms_per_step = ms / args.steps
    value = cells * args.steps / (ms * 1e-3)
    agent_rate = n * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (evaporate+diffuse), live device time per launch
    peak, peak_src = measured_peaks()
    st_ms = prof.stencil_ms / max(1, prof.ticks)
    ach = cells * BYTES_PER_CELL_UPDATE / (st_ms * 1e-3) / 1e9
    tr = stencil_traffic_from_profile()
    if tr and (tr.get("cells") != cells or tr.get("stencil") != cfg.stencil):
        tr = None  # the committed ncu capture is for another shape
    roofline = {"bound": "hbm", "kernel": "evaporate+diffuse stencil", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                "frac_of_8TBs_nominal": ach / 8000.0,
                "traffic": (tr or {}).get("dram_bytes_per_launch"),
                "traffic_source": (None if not tr else f"ncu --set full capture of this kernel, {tr.get('source', 'profiles/')}"
                                                      f" (not measured in this run)"),
                "algorithmic_bytes_per_launch": cells * BYTES_PER_CELL_UPDATE,
                "launch_ms": st_ms,
                "phase_ms": {"deposit": prof.deposit_ms / max(1, prof.ticks), "stencil": st_ms,
                             "agents": prof.agents_ms / max(1, prof.ticks)},
                "whole_tick": {"algorithmic_bytes": cells * 8 + n * (BYTES_PER_AGENT_STEP + 12) + 8 * min(n, cells),
                               "achieved_GBs": (cells * 8 + n * (BYTES_PER_AGENT_STEP + 12) + 8 * min(n, cells))
                               / (ms_per_step * 1e-3) / 1e9}}
    roofline["whole_tick"]["frac"] = roofline["whole_tick"]["achieved_GBs"] / peak

    # ---- end to end through the reference-facing call, host poses in / wheel speeds out
    rc = ReferenceCompat.from_field(field, swarm)
    pose = rc._h_pose
    cur = swarm.to_numpy()
    pose[0].copy_(torch.from_numpy(cur["x"]))
    pose[1].copy_(torch.from_numpy(cur["y"]))
    pose[2].copy_(torch.from_numpy(cur["theta"]))
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    # the "physics" between ticks is not ours to time: two pre-made pinned pose sets 3 mm apart are
    # fed alternately, so every robot has moved (>= 1.5 mm) and deposits on every tick
    pose_b = torch.empty_like(pose).pin_memory()
    pose_b.copy_(pose)
    pose_b[0].add_(3e-3)
    poses = (pose_b, pose)
    for k in range(2):
        rc.tick_host(poses[k % 2])
    torch.cuda.synchronize()
    dep0 = field.counters().deposits
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        wl, wr, dec = rc.tick_host(poses[k % 2])  # synchronous: the supervisor needs the speeds
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": cells * e2e_steps / e2e_s, "unit": "cell-updates/s",
           "agent_steps_per_s": n * e2e_steps / e2e_s, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "steps": e2e_steps, "h2d_bytes_per_step": rc.h2d_bytes, "d2h_bytes_per_step": rc.d2h_bytes,
           "deposits_per_step": (field.counters().deposits - dep0) / e2e_steps,
           "call": "ReferenceCompat.tick_host(pinned poses [3][N]) -> (wl, wr, decision) on host"}

    # ---- the other BASELINE configs and the sparse-trail engine, short runs (context, not the headline)
    others = None
    trail_eng = None
    if args.config == "c3" and not args.no_others and not args.grid:
        others = {}
        for name, k in (("c1", 2000), ("c2", 256), ("c4", 64)):
            try:
                others[name] = measure_other_config(name, k)
            except Exception as e:  # never lose the headline over a side measurement
                others[name] = {"error": repr(e)}
        try:
            trail_eng = measure_trail_engine()
        except Exception as e:
            trail_eng = {"error": repr(e)}

    # ---- CPU baseline (rank 0, N = 1): the oracle port on this box's host cores
    cb = None
    if not args.no_cpu:
        cb, _ = cpu_port_throughput(cfg, n, budget_s=args.cpu_seconds)
        cb["reference_python_legs"] = cpu_reference_legs()

    line = {
        "metric": "grid cell-updates/s (agent-steps/s alongside)", "value": value,
        "unit": "cell-updates/s", "agent_steps_per_s": agent_rate, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_dict(args, cfg, n), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cb,
        "stencil_variant": args.variant, "agent_sort_every": args.sort_every,
        "agent_sorts_in_timed_region": sorts_timed,
        "fused_deposit": {"ticks_attempted": tile_att, "ticks_fell_back": tile_fb, "dense_tiles": tile_dense[0],
                          "dense_tile_buffers": tile_dense[1]},
        "other_configs": others, "sparse_trail_engine": trail_eng,
    }
    emit(line)
    field.close()


def sharded_side_run(name, steps, warmup, rank, world, local, sort_every=32):
    """Another BASELINE config on the same ranks, next to the headline of a sharded run: config 4 (two-channel foraging
    field, clustered population, FSM) on row strips of equal COST, with the per-rank population, strip heights, phase
    times and foraging counters (BASELINE.json configs[4] names 8 B200: the driver's 8-GPU scaling run carries it).
    Every rank always reaches the same collectives (a rank that fails says so in them), so a failure here costs the
    side measurement, never the headline."""
    import torch
    import torch.distributed as dist

    from swarm_robotics_pheromones_b200 import sharded

    def all_ok(ok):
        t = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    strip, err = None, None
    try:
        cfg, n = make_config(name)
        x, y, th, st0 = init_agents_torch(cfg, n, "cuda")
        rows = sharded.agent_rows(cfg, x)
        per_row = np.bincount(rows, minlength=cfg.H).astype(np.float64)
        bounds = None
        if cfg.fsm:  # strips of equal cost: bytes per row = its cells + 260 B (measured DRAM bytes) per robot on it
            bounds = sharded.weighted_bounds(cfg.W * cfg.C * BYTES_PER_CELL_UPDATE + per_row * 260.0, world, min_rows=256)
        scfg = sharded.strip_config(cfg, world, rank, device=local, bounds=bounds)
        mine = np.nonzero((rows >= scfg.row0) & (rows < scfg.row1))[0]
        mcap = max(4096, int(4 * float(per_row.max())) + 4096)
        strip = sharded.CudaStrip(scfg, x[mine], y[mine], th[mine], st0[mine], mine.astype(np.uint32),
                                  int(len(mine) * 1.25) + 4096, mcap)
        del x, y, th, rows
    except Exception as e:
        err = repr(e)
    if not all_ok(err is None):
        if strip is not None:
            strip.field.close()
        return {"error": err or "another rank could not set up its strip"}
    ms = 0.0
    phases = None
    try:
        strip.peer_attach_ipc(rank, world)   # collective; every rank is here
        dist.barrier()
        strip.peer_prime()
        strip.peer_wait()
        drv = sharded.DistDriver(sharded.StripWorker(strip, rank, world))
        cells = cfg.H * cfg.W * cfg.C
        sharded.run_ticks(drv, warmup, sort_every, strip.sort_agents)
        torch.cuda.synchronize()
    except Exception as e:
        err = repr(e)
    dist.barrier()
    if err is None:
        try:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sharded.run_ticks(drv, steps, sort_every, strip.sort_agents)   # the timed region pays ceil(K / sort_every) sorts
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            strip.field.peer_status()
        except Exception as e:
            err = repr(e)
    dist.barrier()
    if err is None:
        try:  # the same ticks once more with CUDA events at the phase boundaries inside pf_strip_tick (untimed)
            strip.field.profile(True)
            sharded.run_ticks(drv, 16, 0, None)
            torch.cuda.synchronize()
            phases = strip.field.strip_profile()
            strip.field.profile(False)
        except Exception as e:
            err = repr(e)
    mine_stats = {"rank": rank, "error": err, "rows": [int(scfg.row0), int(scfg.row1)], "ms": ms, "phases_ms": phases}
    try:
        if err is None:
            c = strip.field.counters()
            mine_stats.update(agents=int(strip.swarm.alive().sum().item()), food_grabbed=int(c.food_grabbed),
                              food_delivered=int(c.food_delivered), deposits=int(c.deposits), ticks=int(c.ticks),
                              migrated_out=int(c.migrated_out), migrate_dropped=int(c.migrate_dropped))
    except Exception as e:
        mine_stats["error"] = repr(e)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine_stats)
    dist.barrier()   # nobody closes its mailboxes while a neighbour may still store into them
    strip.field.close()
    errs = [m["error"] for m in everyone if m.get("error")]
    if errs:
        return {"error": errs[0]}
    ms = max(m["ms"] for m in everyone)
    agents = [m["agents"] for m in everyone]
    out = {"workload": f"{name}: {cfg.H}x{cfg.W}x{cfg.C}, {n} agents, {cfg.stencil}-point, row strips x{world}",
           "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "cell_updates_per_s": cells * steps / (ms * 1e-3),
           "agent_steps_per_s": n * steps / (ms * 1e-3),
           "strips": "equal cost (cells + robots per row)" if bounds is not None else "equal height",
           "strip_rows_per_rank": [m["rows"] for m in everyone], "agents_per_rank": agents,
           "agent_imbalance_max_over_mean": float(max(agents) / max(1.0, float(np.mean(agents)))),
           "migrated_agents_total": sum(m["migrated_out"] for m in everyone),
           "migrate_dropped": sum(m["migrate_dropped"] for m in everyone),
           "phase_ms_per_rank": [m["phases_ms"] for m in everyone],
           "tick_driver": "pf_strip_tick: one C call per tick, halo rows and leavers through peer mailboxes"}
    if cfg.fsm:
        from swarm_robotics_pheromones_b200.metrics import NetEnergy
        delivered, ticks = sum(m["food_delivered"] for m in everyone), everyone[0]["ticks"]
        out["foraging"] = {"food_grabbed": sum(m["food_grabbed"] for m in everyone), "food_delivered": delivered,
                           "deposits": sum(m["deposits"] for m in everyone), "ticks": ticks,
                           "net_energy": NetEnergy().update(delivered, ticks, n)}
    return out


def run_sharded(args):
    import torch
    import torch.distributed as dist

    from swarm_robotics_pheromones_b200 import sharded

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    same_device = os.environ.get("PF_BENCH_SAME_DEVICE") == "1"   # test mode: every rank on cuda:0, gloo for the set-up
    if same_device:
        local = 0
    torch.cuda.set_device(local)
    # stdout carries the one JSON line, so NCCL's log (INFO unless the caller chose a level) goes to a per-rank
    # file and its communicator lines (nranks, transport, NVLS) are copied to STDERR at the end of the run
    os.environ.setdefault("NCCL_DEBUG", "INFO")
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
    nccl_log = None
    if "NCCL_DEBUG_FILE" not in os.environ:
        nccl_log = f"/tmp/pf_nccl_{os.getpid()}.log"
        os.environ["NCCL_DEBUG_FILE"] = nccl_log
    if same_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, n = make_config(args.config, args.grid, args.grid, args.channels, args.agents, args.stencil)
    # every rank draws the same global population and keeps its own strip's agents
    x, y, th, st0 = init_agents_torch(cfg, n, "cuda")
    rows = sharded.agent_rows(cfg, x)
    bounds = None
    balance = args.balance == "on" or (args.balance == "auto" and cfg.fsm)
    if balance:  # strips of equal cost instead of equal height: bytes per row = its cells + the robots on it
        per_row = np.bincount(rows, minlength=cfg.H).astype(np.float64)
        row_cost = cfg.W * cfg.C * BYTES_PER_CELL_UPDATE + per_row * 260.0   # 260 B: measured DRAM bytes per agent-step
        bounds = sharded.weighted_bounds(row_cost, world, min_rows=256)
    scfg = sharded.strip_config(cfg, world, rank, device=local, bounds=bounds)
    mine = np.nonzero((rows >= scfg.row0) & (rows < scfg.row1))[0]
    n_local = len(mine)
    cap = int(n_local * 1.25) + 4096
    per_row_max = float(np.bincount(rows, minlength=cfg.H).max())
    mcap = max(4096, int(4 * per_row_max) + 4096)   # robots that can cross one strip edge in a tick, with slack
    strip = sharded.CudaStrip(scfg, x[mine], y[mine], th[mine], st0[mine], mine.astype(np.uint32), cap, mcap)
    if args.force_owner_deposit:
        strip.field.set_option(abi.OPT_FORCE_OWNER_DEPOSIT, 1)
    if args.strip_one_launch >= 0:
        strip.field.set_option(abi.OPT_STRIP_ONE_LAUNCH, args.strip_one_launch)
    if args.part_a_tiles:
        strip.field.set_option(abi.OPT_STRIP_PART_A_TILES, args.part_a_tiles)
    del x, y, th, rows
    strip.field.set_stencil_variant(args.variant, args.rows_per_cta)
    if same_device:
        strip.field.set_option(abi.OPT_PEER_TIMEOUT_MS, 30000)  # the ranks time-slice one GPU
    worker = sharded.StripWorker(strip, rank, world)
    legacy = args.legacy_loop or args.nccl_halo
    drv = sharded.DistDriver(worker, overlap=not args.no_overlap, profile=args.phase_profile and legacy, legacy=legacy)
    if not args.nccl_halo:
        # halo exchange fused into the deposit / stencil kernels: edge rows are stored straight into
        # the neighbours' ghost rows over NVLink (CUDA IPC mappings); NCCL only carries migrants
        strip.peer_attach_ipc(rank, world)
        dist.barrier()
        strip.peer_prime()
        strip.peer_wait()
    cells = cfg.H * cfg.W * cfg.C

    tick_no = [0]

    def tick(last=False):
        if args.sort_every > 0 and tick_no[0] % args.sort_every == 0:
            strip.sort_agents()
        # the interior rows' next deposits ride on this tick's stencil (pf_tile_rows) — except on the tick
        # before a re-sort and on the last tick of a run (the field would hold deposits of a tick not run)
        before_sort = args.sort_every > 0 and (tick_no[0] + 1) % args.sort_every == 0
        drv.tick(True, hand_over=not (last or before_sort or args.no_hand_over))
        tick_no[0] += 1

    sampler = ClockSampler(local)   # (NVML initialised before the warm-up: no idle gap in front of the timed region)
    for k in range(args.warmup):
        tick(last=k == args.warmup - 1)
    torch.cuda.synchronize()
    dist.barrier()
    tick_no[0] = 0  # the timed region pays ceil(K / sort_every) sorts
    if args.phase_profile and not legacy:
        strip.field.profile(True)   # CUDA events at the phase boundaries inside pf_strip_tick
    l0 = strip.field.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    torch.cuda.synchronize()
    tick_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)] if args.tick_times else None
    e0.record()
    h0 = time.perf_counter()
    for k in range(args.steps):
        if tick_ev:
            tick_ev[k].record()
        tick(last=k == args.steps - 1)
    host_enqueue_ms = (time.perf_counter() - h0) * 1e3
    e1.record()
    if tick_ev:
        tick_ev[args.steps].record()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = strip.field.launch_count() - l0
    mine_ph = None
    if args.phase_profile and not legacy:
        mine_ph = strip.field.strip_profile()
        strip.field.profile(False)
    strip.field.peer_status()  # raises if a bounded peer wait gave up (a rank out of lockstep)
    c_run = strip.field.counters()
    lost = torch.tensor([float(c_run.migrate_dropped), float(c_run.migrated_out), float(c_run.migrated_in)],
                        device="cuda", dtype=torch.float64)
    dist.all_reduce(lost)
    assert lost[0].item() == 0, f"{int(lost[0].item())} migrating agents were dropped (migration mailbox too small)"

    # stencil-only time of this strip (roofline per GPU), same event method
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    torch.cuda.synchronize()
    s0.record()
    for _ in range(reps):
        strip.field.step_field_rows(0, strip.rows)
    s1.record()
    torch.cuda.synchronize()
    st_ms = s0.elapsed_time(s1) / reps
    peak, peak_src = measured_peaks()
    local_cells = strip.rows * cfg.W * cfg.C
    ach = local_cells * BYTES_PER_CELL_UPDATE / (st_ms * 1e-3) / 1e9

    # e2e: per-rank host poses in, wheel speeds out, no motion on the device (sharded.StripHostTick: the strip's
    # counterpart of ReferenceCompat.tick_host)
    sw = strip.swarm
    nl = sw.n
    host = sharded.StripHostTick(strip, worker)
    h_pose = torch.empty((3, nl), dtype=torch.float32).pin_memory()
    h_pose[0].copy_(sw.x.cpu()); h_pose[1].copy_(sw.y.cpu()); h_pose[2].copy_(sw.theta.cpu())
    h_pose_b = torch.empty_like(h_pose).pin_memory()
    h_pose_b.copy_(h_pose)
    h_pose_b[1].add_(3e-3)  # along y: no agent changes strip, every agent has moved
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_k = [0]

    def e2e_tick():
        src = h_pose_b if e2e_k[0] % 2 == 0 else h_pose
        e2e_k[0] += 1
        host.tick(src)   # returns when this rank's speeds and decisions are on the host

    e2e_tick()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_tick()
    torch.cuda.synchronize()
    dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    bytes_io = torch.tensor([host.h2d_bytes, host.d2h_bytes], device="cuda", dtype=torch.float64)
    dist.all_reduce(bytes_io)
    counts = torch.tensor([float(sw.alive().sum().item())], device="cuda", dtype=torch.float64)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    c_end = strip.field.counters()
    food = torch.tensor([float(c_end.food_grabbed), float(c_end.food_delivered), float(c_end.deposits)],
                        device="cuda", dtype=torch.float64)
    dist.all_reduce(food)
    rows_all = [None] * world
    dist.all_gather_object(rows_all, [int(scfg.row0), int(scfg.row1)])

    phases_all = None
    if args.phase_profile:
        phases_all = [None] * world
        if legacy:
            mine_ph = sharded.phase_means(drv)
        dist.all_gather_object(phases_all, mine_ph)
    if rank == 0:
        line = {
            "metric": "grid cell-updates/s (agent-steps/s alongside)",
            "value": cells * args.steps / (ms * 1e-3), "unit": "cell-updates/s",
            "agent_steps_per_s": n * args.steps / (ms * 1e-3), "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_dict(args, cfg, n), "clocks": clocks,
            "e2e": {"value": cells * e2e_steps / e2e_s, "unit": "cell-updates/s",
                    "agent_steps_per_s": n * e2e_steps / e2e_s, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "steps": e2e_steps, "h2d_bytes_per_step": int(bytes_io[0].item()),
                    "d2h_bytes_per_step": int(bytes_io[1].item()),
                    "call": "per rank: sharded.StripHostTick.tick(pinned poses [3][slots]) -> (wl, wr, decision) on host; "
                            "pf_strip_tick phases A+B, D2H beside phase C"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "evaporate+diffuse stencil (rank 0 strip, timed alone)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "peak_source": peak_src, "traffic": None, "launch_ms": st_ms,
                         "algorithmic_bytes_per_launch": local_cells * BYTES_PER_CELL_UPDATE},
            "cpu_baseline": None,
            "agents_per_rank": [int(c.item()) for c in allc],
            "strip_rows_per_rank": rows_all, "strips": "equal cost (cells + robots per row)" if balance else "equal height",
            "agent_imbalance_max_over_mean": float(max(c.item() for c in allc) / max(1.0, np.mean([c.item() for c in allc]))),
            "migrated_agents_total": int(lost[1].item()), "migrate_dropped": int(lost[0].item()),
            "tick_driver": ("host loop, NCCL migration (round-1 path)" if legacy
                            else "pf_strip_tick: one C call per tick, leavers through peer mailboxes"),
            "halo_overlap": not args.no_overlap, "stencil_variant": args.variant,
            "halo": "nccl send/recv" if args.nccl_halo else "peer stores fused into the kernels (CUDA IPC, NVLink P2P)",
            "agent_sort_every": args.sort_every,
            "host_enqueue_ms_per_step_rank0": host_enqueue_ms / args.steps,
        }
        if tick_ev:
            line["tick_ms_rank0"] = [round(tick_ev[k].elapsed_time(tick_ev[k + 1]), 4) for k in range(args.steps)]
        if cfg.fsm:  # the dissertation's outcome metric next to the throughput (PDF p.19-21; metrics.py)
            from swarm_robotics_pheromones_b200.metrics import NetEnergy
            line["foraging"] = {"food_grabbed": int(food[0].item()), "food_delivered": int(food[1].item()),
                                "deposits": int(food[2].item()), "ticks": int(c_end.ticks),
                                "net_energy": NetEnergy().update(int(food[1].item()), int(c_end.ticks), n)}
        if args.phase_profile:
            line["phase_ms_rank0"] = phases_all[0]
            line["phase_ms_all_ranks"] = phases_all
    dist.barrier()
    strip.field.close()
    # config 4 (BASELINE.json: "... foraging field, 4M agents with nest/food sources, 8 B200") on the same ranks
    others = None
    if args.config == "c3" and not args.no_others and (not args.grid or same_device):
        try:
            others = {"c4": sharded_side_run("c4", 32, 5, rank, world, local)}
        except Exception as e:  # collectives out of step would hang, not raise; this is for everything else
            others = {"c4": {"error": repr(e)}}
    if rank == 0:
        line["other_configs"] = others
        emit(line)
    dist.barrier()
    dist.destroy_process_group()
    if nccl_log and os.path.exists(nccl_log):  # the communicator lines, for whoever checks the rank count
        keep = ("nranks", "Init COMPLETE", "NVLS", "comm 0x", "NCCL version", "Connected", "P2P")
        try:
            with open(nccl_log) as fh:
                lines = [ln.rstrip() for ln in fh if any(k in ln for k in keep)]
            for ln in lines[:40] if rank == 0 else lines[:4]:
                print(ln, file=sys.stderr)
            sys.stderr.flush()
            os.unlink(nccl_log)
        except OSError:
            pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--grid", type=int, default=None, help="override H = W (debug)")
    ap.add_argument("--channels", type=int, default=None)
    ap.add_argument("--agents", type=int, default=None)
    ap.add_argument("--stencil", type=int, default=None, choices=[0, 5, 9])
    ap.add_argument("--variant", type=int, default=0, help="stencil kernel variant (0 auto)")
    ap.add_argument("--rows-per-cta", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--sort-every", type=int, default=32,
                    help="re-order the agent arrays by cell every this many ticks (0 = never)")
    ap.add_argument("--overlap-agents", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the c1/c2/c4 and trail side runs")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--nccl-halo", action="store_true", help="sharded run: NCCL send/recv halo instead of peer stores")
    ap.add_argument("--phase-profile", action="store_true", help="sharded run: per-phase CUDA events")
    ap.add_argument("--balance", default="auto", choices=["auto", "on", "off"],
                    help="sharded run: strips of equal cost (cells + robots per row) instead of equal height; "
                         "auto = for config 4's clustered population")
    ap.add_argument("--force-owner-deposit", action="store_true",
                    help="sort-free deposit even where the density heuristic prefers the global radix sort")
    ap.add_argument("--strip-one-launch", type=int, default=-1, help="sharded run: PF_OPT_STRIP_ONE_LAUNCH (0, 1, 2); -1 = default")
    ap.add_argument("--part-a-tiles", type=int, default=0, help="sharded run: PF_OPT_STRIP_PART_A_TILES")
    ap.add_argument("--tick-times", action="store_true", help="sharded run: device time of every timed tick (rank 0)")
    ap.add_argument("--legacy-loop", action="store_true",
                    help="sharded run: round-1 tick driver (phase-by-phase host loop, NCCL migration)")
    ap.add_argument("--no-hand-over", action="store_true",
                    help="sharded run: always use the separate deposit kernel (no deposits riding on the stencil)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner, for one) is
    # sent to stderr instead, and the line itself goes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        args.gpus = world
        run_sharded(args)
    else:
        run_single(args)


if __name__ == "__main__":
    main()
