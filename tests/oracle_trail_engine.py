"""Test-only trail engine with TrailSwarm's interface (startPheromone / set_state / pheromone_grid) over the CPU
trail oracle, so that the headless drop-in and main loop (swarm_robotics_pheromones_b200/headless.py) can be
exercised without a GPU."""
from oracle import trail


class OracleTrailEngine:
    def __init__(self, robot_names, start_positions, food_position, ps=60.0, start_delay=5.0):
        self.names = list(robot_names)
        self.start_delay = start_delay
        self.o = {n: trail.TrailOracle(list(s), list(food_position), ps=(ps,) * 8)
                  for n, s in zip(self.names, start_positions)}

    def set_state(self, name, homing):
        self.o[name].state = trail.HOMING if homing else trail.SEARCHING

    def startPheromone(self, robot_names=None, positions=None, t_now=None):
        out = {}
        for n in (self.names if robot_names is None else robot_names):
            o = self.o[n]
            if o.state == trail.HOMING or t_now >= self.start_delay:
                rec = o.tick(t_now, list(positions[n]))
                if rec["sense"] is not None:
                    rec["sense"] = {k: [[t, i] for t, i in v] for k, v in rec["sense"].items()}
            else:
                rec = {"t": t_now, "amount": None, "n_global": o.n_global, "sense": None, "direction": None,
                       "speeds": None, "n_live": len(o.live), "sum_live": 0.0}
            out[n] = rec
        return out

    homing = startPheromone

    def pheromone_grid(self, name):
        return {k: (v[0], v[1]) for k, v in self.o[name].live.items()}
