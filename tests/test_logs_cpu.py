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


def test_net_energy_formula():
    # E(k+1) = E(k) + F*E_f - E_c (dissertation p.19-21; no reference code: formula check only)
    from swarm_robotics_pheromones_b200.metrics import NetEnergy
    e = NetEnergy(energy_per_food=2000.0, cost_per_robot_step=1.0)
    assert e.update(0, 10, robots=3) == -30.0
    assert e.update(2, 10, robots=3) == 3970.0          # two items delivered, no further ticks
    assert e.update(2, 110, robots=3) == 3970.0 - 300.0
    assert e.history == [-30.0, 3970.0, 3670.0]
    import pytest
    with pytest.raises(ValueError):
        e.update(1, 110, robots=3)
