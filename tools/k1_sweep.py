"""Sweep the evaporate+diffuse kernel variants on a large field (run on the GPU box)."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from swarm_robotics_pheromones_b200 import abi  # noqa: E402
from swarm_robotics_pheromones_b200.field import PheroField  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=32768)
ap.add_argument("--channels", type=int, default=2)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--out", default="gpurun_out/k1_sweep.json")
args = ap.parse_args()

res = []
for stencil in (5, 9, 0):
    cfg = abi.default_config(args.grid, args.grid, args.channels)
    cfg.stencil = stencil
    for c in range(args.channels):
        cfg.diffusion[c] = 0.1
    f = PheroField(cfg)
    f.view().uniform_(0, 1)
    cells = args.grid * args.grid * args.channels
    combos = [(1, r) for r in (16, 32, 64, 128, 256)] + [(v, r) for v in (2, 3, 4, 5) for r in (64, 128, 256, 512)]
    for variant, rpc in combos:
        f.set_stencil_variant(variant, rpc)
        f.step_field(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f.step_field(args.steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        gbs = cells * 8 / ms / 1e6
        res.append({"stencil": stencil, "variant": variant, "rows_per_cta": rpc, "ms": ms, "GBs": gbs})
        print(f"stencil {stencil} variant {variant} rpc {rpc:4d}: {ms:7.3f} ms  {gbs:7.1f} GB/s", flush=True)
    f.close()
# plain copy for reference (same bytes)
a = torch.empty(args.grid * args.grid * args.channels, device="cuda")
b = torch.empty_like(a)
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    b.copy_(a)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(f"torch copy_ same bytes: {ms:7.3f} ms  {a.numel() * 8 / ms / 1e6:7.1f} GB/s")
res.append({"copy_ms": ms, "copy_GBs": a.numel() * 8 / ms / 1e6})
Path(args.out).parent.mkdir(exist_ok=True)
Path(args.out).write_text(json.dumps(res, indent=1))
