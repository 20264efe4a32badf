"""Where does the end-to-end tick go? PCIe rates and the phases of ReferenceCompat.tick_host."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from swarm_robotics_pheromones_b200.compat import ReferenceCompat  # noqa: E402
from swarm_robotics_pheromones_b200.field import PheroField, Swarm  # noqa: E402

n = 1 << 24
h = torch.empty((3, n), dtype=torch.float32).pin_memory()
d = torch.empty((3, n), dtype=torch.float32, device="cuda")
for name, fn in (("H2D 201 MB", lambda: d.copy_(h, non_blocking=True)), ("D2H 201 MB", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt*1e3:.2f} ms  {h.numel()*4/dt/1e9:.1f} GB/s")

cfg, n = bench.make_config("c3")
x, y, th = bench.init_agents_torch(cfg, n, "cuda")
f = PheroField(cfg)
sw = Swarm(x, y, th)
f.sort_agents(sw)
rc = ReferenceCompat.from_field(f, sw)
pose = rc._h_pose
cur = sw.to_numpy()
pose[0].copy_(torch.from_numpy(cur["x"])); pose[1].copy_(torch.from_numpy(cur["y"])); pose[2].copy_(torch.from_numpy(cur["theta"]))
for _ in range(3):
    rc.tick_host(pose)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
main = torch.cuda.current_stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
ev[0].record()
rc._d_pose.copy_(pose, non_blocking=True)
ev[1].record()
f.observe(sw, rc._d_pose[0], rc._d_pose[1], rc._d_pose[2])
f.agents_deposit(sw)
ev[2].record()
f.agents_step(sw, 2)
ev[3].record()
rc._h_out[0].copy_(sw.wl, non_blocking=True)
rc._h_out[1].copy_(sw.wr, non_blocking=True)
ev[4].record()
f.step(sw, 1, 4)
ev[5].record()
torch.cuda.synchronize()
t1 = time.perf_counter()
names = ["H2D", "observe+deposit", "agents", "D2H (serial here)", "stencil"]
for i, nm in enumerate(names):
    print(f"{nm}: {ev[i].elapsed_time(ev[i+1]):.2f} ms")
print(f"serial total {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter()
for _ in range(5):
    rc.tick_host(pose)
print(f"tick_host: {(time.perf_counter()-t0)/5*1e3:.2f} ms")
