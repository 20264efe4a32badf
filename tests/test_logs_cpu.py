"""heatmap.txt round trip in the reference's format (sup:697-701 writer, graph.py:13-15 reader)."""
import json

from swarm_robotics_pheromones_b200.logs import TrajectoryLog


def test_heatmap_txt_round_trip(tmp_path, golden_dir):
    pts = json.load(open(golden_dir / "sup_replay_heatmap.json"))["positions"]
    log = TrajectoryLog(str(tmp_path / "heatmap.txt"))
    for p in pts:
        log.append(p)
    log.flush()
    back = TrajectoryLog.read(str(tmp_path / "heatmap.txt"))
    assert back == pts  # repr round-trips doubles exactly, as in the reference's own file
    text = open(tmp_path / "heatmap.txt").read()
    assert text.count("\n") == 1 and text.startswith("[[0.35026199999999985, -0.4114980000683755")


def test_rewrite_every_tick_and_2d_positions(tmp_path):
    log = TrajectoryLog(str(tmp_path / "h.txt"), rewrite_every_tick=True)
    log.append([0.1, 0.2])
    log.append([0.3, 0.4, 0.5])
    assert TrajectoryLog.read(str(tmp_path / "h.txt")) == [[0.1, 0.2, 0.0], [0.3, 0.4, 0.5]]
