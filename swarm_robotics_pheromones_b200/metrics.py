"""Foraging outcome metrics from the device counters.

The dissertation defines the swarm's net energy (PDF p.19-21) as

    E(k+1) = E(k) + F(k) * E_f - E_c(k)

with F(k) the food items delivered to the nest in step k, E_f the energy of one item and E_c the
energy the swarm spends in step k. The reference code never implements it (it only keeps
`food_counter`, experiments.py:554,734), so there is no reference arithmetic to be equal
to: this is the formula as printed, evaluated in float64 on the host from `pf_counters`
(food_delivered) and the number of live robots. PARITY UNPINNED.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List


@dataclass
class NetEnergy:
    energy_per_food: float = 2000.0      # E_f
    cost_per_robot_step: float = 1.0     # E_c(k) = cost_per_robot_step * robots active in step k
    energy: float = 0.0                  # E(k)
    delivered_seen: int = 0
    ticks_seen: int = 0
    history: List[float] = field(default_factory=list)

    def update(self, food_delivered_total: int, ticks_total: int, robots: int) -> float:
        """advance to the counters' current tick; call as often as the curve should be sampled.
        All ticks since the last call are charged E_c, all deliveries since then are credited."""
        df = int(food_delivered_total) - self.delivered_seen
        dt = int(ticks_total) - self.ticks_seen
        if df < 0 or dt < 0:
            raise ValueError("counters went backwards (reset?): start a new NetEnergy")
        self.energy += df * self.energy_per_food - dt * robots * self.cost_per_robot_step
        self.delivered_seen += df
        self.ticks_seen += dt
        self.history.append(self.energy)
        return self.energy

    def update_from(self, field_handle, robots: int) -> float:
        """`field_handle`: a PheroField (reads pf_counters_host)"""
        c = field_handle.counters()
        return self.update(c.food_delivered, c.ticks, robots)
