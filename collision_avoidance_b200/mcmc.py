"""Batched MCMC search over ALAN action sets.

Mirrors collision_avoidance/ALAN/Train_ALAN_action_space.py (``MCMC_trainer``: ``train`` :27,
``evaluate_action`` :55, ``select_modification`` :70, ``mod_edit/remove/add`` :86-126,
``symmetric_likelihood`` :133) with the same method names.  The reference evaluates ONE
candidate action set with 3 sequential simulations per round; here ``chains`` independent
annealing chains run side by side and every round evaluates all their candidates in ONE batch
of ``chains * sims_per_eval`` worlds, each world carrying its own action table (per-world
tables are a kernel feature: OrcaEnvStepArgs.alan_actions_env_stride).

Deliberate differences from the reference, which has two bugs there:
  * candidate sets are COPIES -- the reference edits ``self.actions`` in place, so its
    accept/reject step never rejects and ``actions_opt`` aliases the working set;
  * the temperature anneals DOWN from 0.9 to 0.1 -- the reference subtracts a negative delta
    and heats up instead.  ``reference_temperature=True`` restores that schedule.
"""
from __future__ import annotations

from math import atan2, cos, exp, pi, sin, sqrt
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .alan import Collision_Avoidance_Sim

Action = Tuple[float, float]


def _normal_pdf(x: float) -> float:
    return exp(-0.5 * x * x) / sqrt(2 * pi)  # scipy.stats.norm.pdf


class MCMC_trainer:
    def __init__(self, numAgents: int = 50, scenario: str = "crowd", numRounds: int = 10, chains: int = 1,
                 sims_per_eval: int = 3, seed: int = 0, device="cuda:0", reference_temperature: bool = False,
                 max_steps: Optional[int] = None):
        self.numAgents, self.scenario, self.numRounds = numAgents, scenario, numRounds
        self.chains, self.sims_per_eval = int(chains), int(sims_per_eval)
        self.rng = np.random.default_rng(seed)
        self.max_steps = max_steps
        self.simulator = Collision_Avoidance_Sim(numAgents=numAgents, scenario=scenario, visualize=False,
                                                 num_envs=self.chains * self.sims_per_eval, seed=seed, device=device)
        # one working set + one best set per chain (Train_ALAN_action_space.py:16-20)
        self.actions: List[List[Action]] = [[(1, 0), self.random_action()] for _ in range(self.chains)]
        self.actions_opt = [list(a) for a in self.actions]
        self.eval = self.evaluate_action(self.actions)
        self.eval_opt = list(self.eval)
        self.init_temp, self.final_temp = 0.9, 0.1
        self.temp = self.init_temp
        self.delta_temp = (self.final_temp - self.init_temp) / max(1, self.numRounds - 1)
        self.reference_temperature = reference_temperature
        self.history = []

    # ------------------------------------------------------------------ search loop
    def train(self):
        for i in range(self.numRounds):
            proposals, dists = [], []
            for c in range(self.chains):
                modification = self.select_modification(self.actions[c], i)
                d, new_actions = self.apply_modification(list(self.actions[c]), modification)
                proposals.append(new_actions)
                dists.append(d)
            new_eval = self.evaluate_action(proposals, i)
            for c in range(self.chains):
                if new_eval[c] < self.eval_opt[c]:
                    self.actions_opt[c], self.eval_opt[c] = list(proposals[c]), new_eval[c]
                accept = self.symmetric_likelihood(dists[c]) * exp(min(50.0, (self.eval[c] - new_eval[c]) / self.temp))
                if self.rng.uniform(0, 1) < accept:
                    self.actions[c], self.eval[c] = proposals[c], new_eval[c]
            self.history.append((i, self.temp, min(self.eval_opt)))
            self.temp = self.temp - self.delta_temp if self.reference_temperature else self.temp + self.delta_temp
        best = int(np.argmin(self.eval_opt))
        return self.actions_opt[best]

    # ------------------------------------------------------------------ pieces (same names as the reference)
    def random_action(self) -> Action:
        angle = self.rng.uniform(-pi, pi)
        return cos(angle), sin(angle)

    def evaluate_action(self, actions: Sequence[Sequence[Action]], i: int = 0) -> List[float]:
        """Mean TTime over ``sims_per_eval`` simulations for every chain's candidate, one batch."""
        per_world = [list(a) for a in actions for _ in range(self.sims_per_eval)]
        self.simulator.reset(online_actions=per_world)
        finished, total_time, ttime, min_ttime = self.simulator.run_sim(mode=1, max_steps=self.max_steps)
        t = ttime.reshape(self.chains, self.sims_per_eval).mean(1)
        return [float(x) for x in t.cpu()]

    def select_modification(self, actions, i) -> int:
        if len(actions) <= 1:
            return 2
        return int(self.rng.choice(3, p=[0.8, 0.1, 0.1]))

    def apply_modification(self, actions, modification):
        if modification == 2 and len(actions) < 16:
            return self.mod_add(actions)
        if modification == 1 and len(actions) > 1:
            return self.mod_remove(actions)
        return self.mod_edit(actions) if len(actions) > 1 else self.mod_add(actions)

    def mod_edit(self, actions):
        index = int(self.rng.integers(1, len(actions)))
        angle = atan2(actions[index][1], actions[index][0])
        new_angle = self.rng.normal(angle, pi)
        new_action = (cos(new_angle), sin(new_angle))
        d = self.dist(actions[index], new_action)
        actions[index] = new_action
        return d, actions

    def mod_remove(self, actions):
        index = int(self.rng.integers(1, len(actions)))
        old = actions.pop(index)
        return min([10.0] + [self.dist(a, old) for a in actions]), actions

    def mod_add(self, actions):
        index = int(self.rng.integers(0, len(actions)))
        angle = atan2(actions[index][1], actions[index][0])
        new_angle = self.rng.normal(angle, pi)
        new_action = (cos(new_angle), sin(new_angle))
        d = self.dist(actions[index], new_action)
        actions.append(new_action)
        return d, actions

    @staticmethod
    def dist(p1, p2) -> float:
        return sqrt((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2)

    @staticmethod
    def symmetric_likelihood(d: float) -> float:
        return _normal_pdf(d) / 0.5
