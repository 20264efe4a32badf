"""Multi-GPU parity under torchrun (NCCL): G row strips on G GPUs vs the unsharded run on rank 0's
GPU, bit for bit (field and agents by gid).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_parity.py --ticks 100 --peer

--peer: halo rows and migrating agents cross as peer stores into CUDA-IPC mailboxes, the whole tick is one
pf_strip_tick call (the product path); --legacy keeps the round-1 host loop with NCCL migration.
--same-device: every rank uses cuda:0 and gloo carries the set-up (NCCL refuses two ranks on one GPU) — the
cross-process mailbox protocol on a box with a single GPU (tests/test_gpu_multiproc.py).
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from swarm_robotics_pheromones_b200 import abi, sharded  # noqa: E402
from swarm_robotics_pheromones_b200.field import PheroField, Swarm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ticks", type=int, default=100)
ap.add_argument("--grid", type=int, default=2048)
ap.add_argument("--agents", type=int, default=200000)
ap.add_argument("--no-overlap", action="store_true")
ap.add_argument("--sort-every", type=int, default=16)
ap.add_argument("--no-hand-over", action="store_true")
ap.add_argument("--peer", action="store_true", help="peer-store halo (CUDA IPC + NVLink P2P) instead of NCCL")
ap.add_argument("--legacy", action="store_true", help="phase-by-phase host loop with NCCL migration (round 1)")
ap.add_argument("--host-ticks", type=int, default=0,
                help="after the ticks: this many host-facing ticks (sharded.StripHostTick: pinned poses in, speeds out, "
                     "no motion on the device) against observe + pf_step_ex without PF_STEP_MOVE on the unsharded field")
ap.add_argument("--same-device", action="store_true", help="all ranks on cuda:0, gloo for the set-up")
args = ap.parse_args()

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = 0 if args.same_device else int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
if args.same_device:
    assert args.peer and not args.legacy, "--same-device needs the mailbox path (--peer, no --legacy)"
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

H = W = args.grid
cfg = abi.default_config(H, W, 2)
cfg.stencil = 9
cfg.fsm = 1
for c in range(2):
    cfg.diffusion[c] = 0.1
    cfg.evap_keep[c] = 0.95
cfg.nest_x, cfg.nest_y = 0.5 * H * 0.01, 0.5 * W * 0.01
cfg.food_x[0], cfg.food_y[0] = 0.7 * H * 0.01, 0.6 * W * 0.01
cfg.food_radius = 0.5
rng = np.random.default_rng(42)
n = args.agents
x = rng.uniform(0.05, H * 0.01 - 0.05, n).astype(np.float32)
y = rng.uniform(0.05, W * 0.01 - 0.05, n).astype(np.float32)
th = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
st = rng.integers(1, 3, n)

parts = sharded.partition_agents(cfg, world, x, y, th, st)
p = parts[rank]
scfg = sharded.strip_config(cfg, world, rank, device=local)
strip = sharded.CudaStrip(scfg, p["x"], p["y"], p["theta"], p["state"], p["gid"],
                          capacity=int(len(p["x"]) * 1.5) + 1024, migrate_cap=8192)
if args.same_device:
    strip.field.set_option(abi.OPT_PEER_TIMEOUT_MS, 20000)  # the ranks time-slice one GPU: waits span context switches
drv = sharded.DistDriver(sharded.StripWorker(strip, rank, world), overlap=not args.no_overlap, legacy=args.legacy)
if args.peer:
    strip.peer_attach_ipc(rank, world)
    dist.barrier()
    strip.peer_prime()
    strip.peer_wait()
for t in range(args.ticks):
    if args.sort_every and t % args.sort_every == 0:
        strip.sort_agents()
    # interior rows hand their next deposit to the stencil, except before a re-sort and on the last tick
    before_sort = args.sort_every and (t + 1) % args.sort_every == 0
    drv.tick(True, hand_over=not (before_sort or t == args.ticks - 1 or args.no_hand_over))
    if args.same_device and t % 8 == 7:
        strip.field.peer_status()  # stop early (raises) instead of piling up timeouts if the ranks cannot overlap
torch.cuda.synchronize()


def host_pose(xv, yv, thv, gid, k):
    """the 'physics' of host tick k: every robot shifted along y by a gid-dependent step (nobody changes strip)"""
    step = ((gid.to(torch.int64) + k) % 5 - 2).to(torch.float32) * 1.5e-3
    ymax = W * 0.01 - 0.02
    return xv, torch.clamp(yv + step, 0.02, ymax), thv


host_ok = True
if args.host_ticks:
    host = sharded.StripHostTick(strip, drv.w)
    sw_ = strip.swarm
    pose = torch.empty((3, sw_.n), dtype=torch.float32).pin_memory()
    for k in range(args.host_ticks):
        nx, ny, nth = host_pose(sw_.x, sw_.y, sw_.theta, sw_.gid, k)
        pose[0].copy_(nx); pose[1].copy_(ny); pose[2].copy_(nth)
        torch.cuda.synchronize()
        wl, wr, dec = host.tick(pose)
        # what arrived on the host is what the device holds
        host_ok = host_ok and bool((wl.view(torch.int32) == sw_.wl.cpu().view(torch.int32)).all())
        host_ok = host_ok and bool((wr.view(torch.int32) == sw_.wr.cpu().view(torch.int32)).all())
        want_dec = ((sw_.meta & abi.META_DEC_MASK) >> abi.META_DEC_SHIFT).to(torch.uint8).cpu()
        host_ok = host_ok and bool((dec == want_dec).all())
    torch.cuda.synchronize()

strip.field.peer_status()  # raises if a bounded peer wait gave up

# gather strips on rank 0
sw = strip.swarm.to_numpy()
alive = (sw["meta"] & 3) != 0
rec = np.stack([sw[f][alive].view(np.uint32) for f in ("gid", "x", "y", "theta", "wl", "wr", "meta")], 1)
gathered = [None] * world if rank == 0 else None
dist.gather_object((strip.field.snapshot(), rec), gathered, dst=0)
ok = True
if rank == 0:
    field = np.concatenate([g_[0] for g_ in gathered], axis=1)
    recs = [g_[1] for g_ in gathered]
    allrec = np.concatenate(recs)
    allrec = allrec[np.argsort(allrec[:, 0], kind="stable")]
    # unsharded run on this GPU
    g = PheroField(cfg)
    s1 = Swarm(x, y, th, st)
    g.step(s1, args.ticks)
    for k in range(args.host_ticks):
        nx, ny, nth = host_pose(s1.x, s1.y, s1.theta, s1.gid, k)
        g.observe(s1, nx.clone(), ny.clone(), nth.clone())
        g.step(s1, 1, abi.STEP_DEPOSIT | abi.STEP_SENSE | abi.STEP_FIELD)
    torch.cuda.synchronize()
    want_f = g.snapshot()
    w = s1.to_numpy()
    g.close()
    same_field = bool((field.view(np.uint32) == want_f.view(np.uint32)).all())
    same_n = allrec.shape[0] == n
    same_agents = same_n and all(
        (allrec[:, k + 1] == w[f].view(np.uint32)).all()
        for k, f in enumerate(("x", "y", "theta", "wl", "wr", "meta")))
    mig = strip.field.counters()
    ok = same_field and same_agents
    if args.host_ticks:
        print(f"dist_parity: {args.host_ticks} host-facing ticks (StripHostTick) included in the comparison", flush=True)
    att, fb = strip.field.tile_stats()
    print(f"dist_parity world={world} hand-over ticks on rank 0: {att} (fell back {fb}); peer_halo={args.peer} path={'legacy host loop + NCCL migration' if (args.legacy or not args.peer) else 'pf_strip_tick + mailbox migration'} ticks={args.ticks} grid={H} agents={n}: field_bitexact={same_field} "
          f"agents_bitexact={same_agents} (rank0 migrated_out={mig.migrated_out}, dropped={mig.migrate_dropped}) "
          f"-> {'PASS' if ok else 'FAIL'}", flush=True)
if args.host_ticks:  # every rank's host buffers held what its device arrays hold
    flag = torch.tensor([1 if host_ok else 0], dtype=torch.int32, device="cuda" if not args.same_device else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not bool(flag.item()):
        ok = False
        if rank == 0:
            print("dist_parity: host buffers differ from the device arrays -> FAIL", flush=True)
dist.barrier()
strip.field.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
