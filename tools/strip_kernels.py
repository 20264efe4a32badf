#!/usr/bin/env python
"""All G strips of the config-3 arena in ONE process on ONE GPU (LocalDriver + pf_strip_tick, peer mailboxes by raw
pointers): the kernels of a strip tick at the sizes they have at G GPUs, for a per-kernel launch list under ncu

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/strip_launches.csv \
        python tools/strip_kernels.py --world 8 --ticks 6

(ncu cannot follow a multi-rank torchrun job; the kernels and their launch dimensions are the same here)."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from swarm_robotics_pheromones_b200 import abi, sharded  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--ticks", type=int, default=6)
ap.add_argument("--config", default="c3")
ap.add_argument("--one-launch", type=int, default=-1, help="PF_OPT_STRIP_ONE_LAUNCH (0, 1, 2); -1 = the library default")
ap.add_argument("--strips", type=int, default=0, help="materialise only the first K strips (0 = all): saves memory")
args = ap.parse_args()

cfg, n = bench.make_config(args.config)
x, y, th, st = bench.init_agents_torch(cfg, n, "cuda")
rows = sharded.agent_rows(cfg, x)
G = args.world
workers = []
for r in range(G):
    scfg = sharded.strip_config(cfg, G, r, device=0)
    mine = np.nonzero((rows >= scfg.row0) & (rows < scfg.row1))[0]
    strip = sharded.CudaStrip(scfg, x[mine], y[mine], th[mine], st[mine], mine.astype(np.uint32),
                              int(len(mine) * 1.25) + 4096, 8192)
    if args.one_launch >= 0:
        strip.field.set_option(abi.OPT_STRIP_ONE_LAUNCH, args.one_launch)
    workers.append(sharded.StripWorker(strip, r, G))
drv = sharded.LocalDriver(workers)
drv.attach_peers()
for w in workers:
    w.b.sort_agents()
torch.cuda.synchronize()
t0 = time.perf_counter()
sharded.run_ticks(drv, args.ticks)
torch.cuda.synchronize()
print(f"{G} strips, {args.ticks} ticks: {(time.perf_counter() - t0) / args.ticks * 1e3:.3f} ms per tick for ALL strips on one GPU")
for w in workers:
    w.b.field.peer_status()
    w.b.field.close()
