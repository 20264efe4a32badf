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
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--config", default="c3")
ap.add_argument("--one-launch", type=int, default=-1, help="PF_OPT_STRIP_ONE_LAUNCH (0, 1, 2); -1 = the library default")
ap.add_argument("--mig-inline", type=int, default=-1, help="PF_OPT_STRIP_MIG_INLINE; -1 = the library default")
ap.add_argument("--no-predeposit", type=int, default=-1, help="PF_OPT_STRIP_NO_PREDEPOSIT; -1 = the library default")
ap.add_argument("--edges-side", type=int, default=-1, help="PF_OPT_STRIP_EDGES_SIDE; -1 = the library default")
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
    if args.no_predeposit >= 0:
        strip.field.set_option(abi.OPT_STRIP_NO_PREDEPOSIT, args.no_predeposit)
    if args.edges_side >= 0:
        strip.field.set_option(abi.OPT_STRIP_EDGES_SIDE, args.edges_side)
    if args.mig_inline >= 0:
        strip.field.set_option(abi.OPT_STRIP_MIG_INLINE, args.mig_inline)
    workers.append(sharded.StripWorker(strip, r, G))
drv = sharded.LocalDriver(workers)
drv.attach_peers()
for w in workers:
    w.b.sort_agents()
sharded.run_ticks(drv, 8)   # warm-up: lazy allocations, clocks
torch.cuda.synchronize()
reps = []
for _ in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sharded.run_ticks(drv, args.ticks)
    e1.record()
    torch.cuda.synchronize()
    reps.append(e0.elapsed_time(e1) / args.ticks)
print(f"{G} strips, {args.reps} x {args.ticks} ticks (the last one of each run without hand-over): "
      f"{' / '.join(f'{r:.3f}' for r in reps)} ms per tick for ALL strips on one GPU (CUDA events), "
      f"best {min(reps) / G:.4f} ms per strip tick")
for w in workers:
    w.b.field.peer_status()
    w.b.field.close()
