"""Generate the committed golden fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py            # needs /root/reference (or PF_REFERENCE_ROOT)

Writes, next to this file:
  sup_replay_heatmap.json  Phase-3/heatmap.txt (760 recorded positions of robot e2 in testing.wbt)
                           replayed through Controller.startPheromone at dt = 0.064 s from t = 5 s,
                           ps_j = 60: per tick deposit amount, global count, sense dict, arg-max
                           direction, wheel speeds, live entries and their sum.
  sup_replay_homing.json   a synthetic approach to the nest replayed through Controller.homing.
  pheromonelevels.json     Phase-3/pheromonelevels.txt (36 intensities) as shipped.
  algo_dense.json          Phase-3/pheromone_algo.py's 10x10 grid after 25 loop iterations.
  graph_heatmap.json       Phase-3/graph.py's deposit + gaussian_filter loop on heatmap.txt.

The GPU box has no reference checkout: tests only read these JSON files.
"""
from __future__ import annotations

import ast
import json
import math
import runpy
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import ref_harness as rh  # noqa: E402

REF = rh.REFERENCE_ROOT
DT = 0.064
T0 = 5.0


def heatmap_points():
    txt = (REF / "Phase-3/heatmap.txt").read_text().strip().splitlines()[-1]
    return ast.literal_eval(txt)


def slim(rec, keep_live):
    out = {k: rec[k] for k in ("t", "amount", "n_global", "sense", "direction", "speeds", "n_live",
                               "sum_live")}
    if keep_live:
        out["live"] = rec["live"]
    return out


def make_sup_replay():
    pts = heatmap_points()
    drv = rh.ReferenceDriver(pts[0], food=(0.36, -0.15, 0.05), ps=60.0)
    ticks = []
    for i, p in enumerate(pts):
        rec = drv.tick(T0 + i * DT, p)
        ticks.append(slim(rec, keep_live=(i % 40 == 0 or i == len(pts) - 1)))
    out = {"source": "Phase-3/heatmap.txt replayed through the unmodified "
                     "Phase-3/pheromone_algo_supervisor.py (Controller.startPheromone)",
           "dt": DT, "t0": T0, "ps": 60.0, "food": [0.36, -0.15, 0.05], "start": pts[0],
           "positions": pts, "ticks": ticks}
    (HERE / "sup_replay_heatmap.json").write_text(json.dumps(out))
    print("sup_replay_heatmap.json", len(ticks), "ticks")


def homing_path():
    """A deterministic walk that visits every HOMING branch of sup:475-514: far (> 0.7 m), inside the
    radius with turn in (-2,0), turn < -2, turn >= 0 (KEEP), then the 5..6 cm ring (STOP)."""
    start = (0.0, 0.0, 0.0)
    pts = []
    # far, moving in
    for k in range(30):
        r = 1.2 - 0.02 * k
        pts.append((r * math.cos(0.3), r * math.sin(0.3), 0.0))
    # circle at 0.5 m through all bearings
    for k in range(120):
        a = 0.3 + 2 * math.pi * k / 120
        pts.append((0.5 * math.cos(a), 0.5 * math.sin(a), 0.0))
    # standing still (no deposit) for 3 ticks
    for _ in range(3):
        pts.append(pts[-1])
    # approach to the 5..6 cm ring
    a = 0.3
    r = 0.5
    while r > 0.056:
        r -= 0.004
        pts.append((r * math.cos(a), r * math.sin(a), 0.0))
    for _ in range(3):
        pts.append(pts[-1])
    return start, pts


def make_homing_replay():
    start, pts = homing_path()
    drv = rh.ReferenceDriver(start, food=(0.36, -0.15, 0.05), ps=60.0)
    drv.set_state(homing=True)
    ticks = []
    for i, p in enumerate(pts):
        rec = drv.tick(T0 + i * DT, list(p))
        ticks.append(slim(rec, keep_live=(i % 40 == 0 or i == len(pts) - 1)))
    out = {"source": "synthetic approach replayed through the unmodified reference "
                     "(Controller.homing, robot_state = HOMING)",
           "dt": DT, "t0": T0, "ps": 60.0, "food": [0.36, -0.15, 0.05], "start": list(start),
           "positions": [list(p) for p in pts], "ticks": ticks}
    (HERE / "sup_replay_homing.json").write_text(json.dumps(out))
    print("sup_replay_homing.json", len(ticks), "ticks")


def make_levels():
    d = ast.literal_eval((REF / "Phase-3/pheromonelevels.txt").read_text().strip())
    (HERE / "pheromonelevels.json").write_text(json.dumps(
        {"source": "Phase-3/pheromonelevels.txt", "e2": d["e2"]}))
    print("pheromonelevels.json", len(d["e2"]))


def make_algo_dense(iterations=25):
    """Run Phase-3/pheromone_algo.py under a fake `controller.Robot`."""
    rh._install_stubs()
    ctrl = sys.modules["controller"]

    class Dev:
        def __init__(self, name):
            self.name = name

        def enable(self, *_):
            pass

        def getValue(self):
            return 4096 * 3 + 1.0 if self.name == "ps0" else (4096 * 7 + 1.0 if self.name == "ps2" else 0.0)

        def set(self, *_):
            pass

        def setPosition(self, *_):
            pass

        def setVelocity(self, *_):
            pass

        def getWidth(self):
            return 4

        def getHeight(self):
            return 4

        def getImage(self):
            return bytes(4 * 4 * 4)

        def getQueueLength(self):
            return 0

        def send(self, *_):
            pass

    class Robot:
        calls = 0

        def getBasicTimeStep(self):
            return 32

        def getDevice(self, name):
            return Dev(name)

        def getNumberOfDevices(self):
            return 0

        def getName(self):
            return "e1"

        def getTime(self):
            return Robot.calls * 0.032

        def step(self, _ts):
            Robot.calls += 1
            # the while-condition call is every odd call; stop after `iterations` loop bodies
            return -1 if Robot.calls == 2 * iterations + 1 else 0

    ctrl.Robot = Robot
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        g = runpy.run_path(str(REF / "Phase-3/pheromone_algo.py"), run_name="__main__")
    grid = g["pheromone_grid"]
    out = {"source": "Phase-3/pheromone_algo.py via runpy with a fake controller.Robot; raw sensor "
                     "ps0 = 3*4096+1, ps2 = 7*4096+1 -> cell (3,7)",
           "iterations": iterations, "cell": [3, 7], "grid": grid,
           "sum": float(np.sum(grid)), "closed_form": 9 * (1 - 0.9 ** iterations)}
    (HERE / "algo_dense.json").write_text(json.dumps(out))
    print("algo_dense.json sum", out["sum"], "closed form", out["closed_form"])


def make_graph():
    """graph.py:17-24 verbatim semantics with scipy (the reference's own dependency)."""
    import cv2
    from scipy.ndimage import gaussian_filter
    bg = cv2.imread(str(REF / "Phase-3/background.png"), cv2.IMREAD_UNCHANGED)
    height, width = bg.shape[:2]
    pts = heatmap_points()
    heatmap = np.zeros((height, width))
    snap = {}
    for i, line in enumerate(pts):
        scale_x = (line[0] / 1) * width
        scale_y = (line[1] / 1) * height
        heatmap[abs(int(scale_y)), abs(int(scale_x))] += 20
        heatmap = gaussian_filter(heatmap, sigma=5)
        if i + 1 == 40:
            r, c = np.unravel_index(np.argmax(heatmap), heatmap.shape)
            snap = {"n": 40, "argmax": [int(r), int(c)], "max": float(heatmap.max()),
                    "sum": float(heatmap.sum()),
                    "crop_origin": [int(r) - 8, int(c) - 8],
                    "crop": heatmap[r - 8:r + 9, c - 8:c + 9].tolist()}
    r, c = np.unravel_index(np.argmax(heatmap), heatmap.shape)
    out = {"source": "Phase-3/graph.py:17-24 on Phase-3/heatmap.txt, scipy.ndimage.gaussian_filter",
           "shape": [int(height), int(width)], "n_points": len(pts),
           "argmax": [int(r), int(c)], "max": float(heatmap.max()), "sum": float(heatmap.sum()),
           "after_40": snap,
           "block_sums_7x5": heatmap[:497 // 71 * 71, :495 // 99 * 99].reshape(7, 71, 5, 99).sum(
               axis=(1, 3)).tolist()}
    (HERE / "graph_heatmap.json").write_text(json.dumps(out))
    print("graph_heatmap.json argmax", out["argmax"], "max", out["max"], "sum", out["sum"])


if __name__ == "__main__":
    if not rh.available():
        sys.exit(f"reference checkout not found at {REF}")
    make_sup_replay()
    make_homing_replay()
    make_levels()
    make_algo_dense()
    make_graph()
