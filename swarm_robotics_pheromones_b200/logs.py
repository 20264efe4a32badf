"""The reference's on-disk trajectory format (Phase-3/heatmap.txt), written from device positions.

`Controller.mappingConcentration` (sup:697-701) rewrites `heatmap.txt` every tick with the repr of the
list of all positions so far: one line, `[[x, y, z], [x, y, z], ...]`. graph.py:13-15 reads the last
line back with `ast.literal_eval`. TrajectoryLog keeps that format for one robot of a Swarm / a
TrailSwarm so the reference's own plotting script (and HeatMap.accumulate) can consume it."""
from __future__ import annotations

from typing import List


class TrajectoryLog:
    def __init__(self, path: str = "heatmap.txt", rewrite_every_tick: bool = False):
        self.path = path
        self.positions: List[List[float]] = []
        self.rewrite_every_tick = rewrite_every_tick  # the reference's O(ticks) per tick behaviour

    def append(self, pos) -> None:
        p = [float(v) for v in pos]
        if len(p) == 2:
            p.append(0.0)
        self.positions.append(p)
        if self.rewrite_every_tick:
            self.flush()

    def append_from_swarm(self, swarm, index: int = 0) -> None:
        """position of slot `index` of a Swarm (x, y on the device; z = 0)"""
        self.append([float(swarm.x[index].item()), float(swarm.y[index].item()), 0.0])

    def flush(self) -> None:
        with open(self.path, "w") as f:  # sup:698-701
            f.write("{}\n".format(self.positions))
            f.flush()

    @staticmethod
    def read(path: str):
        """graph.py:13-15: the last line, literal-eval'd"""
        import ast
        sub = None
        with open(path) as f:
            for line in f:
                sub = ast.literal_eval(line)
        return sub
