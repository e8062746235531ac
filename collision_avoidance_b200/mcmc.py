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
``reference_semantics=True`` switches both back (in-place edits, aliasing, rising temperature):
with it, chain ``c`` reproduces the reference's ``train()`` move for move after
``np.random.seed(seed + c); random.seed(seed + c)`` -- the random draws below follow the
reference's call pattern on the same generators (legacy ``RandomState.choice`` / ``normal``,
``random.uniform``).  Pinned against the unmodified trainer in tests/golden/shell_mcmc.json
(tests/test_shell_golden.py).
"""
from __future__ import annotations

import random as _pyrandom
from math import cos, exp, pi, sin, sqrt
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .alan import Collision_Avoidance_Sim

Action = Tuple[float, float]


def _normal_pdf(x: float) -> float:
    return exp(-0.5 * x * x) / sqrt(2 * pi)  # scipy.stats.norm.pdf


class MCMC_trainer:
    def __init__(self, numAgents: int = 50, scenario: str = "crowd", numRounds: int = 10, chains: int = 1,
                 sims_per_eval: int = 3, seed: int = 0, device="cuda:0", reference_temperature: bool = False,
                 max_steps: Optional[int] = None, reference_semantics: bool = False,
                 cost_fn: Optional[Callable[[Sequence[Action]], float]] = None):
        self.numAgents, self.scenario, self.numRounds = numAgents, scenario, numRounds
        self.chains, self.sims_per_eval = int(chains), int(sims_per_eval)
        # the reference draws from the module-level numpy (legacy) and Python generators
        # (Train_ALAN_action_space.py:41,52,75,89,92,...); one pair of those streams per chain
        self.np_rngs = [np.random.RandomState(seed + c) for c in range(self.chains)]
        self.py_rngs = [_pyrandom.Random(seed + c) for c in range(self.chains)]
        self._chain = 0
        self.max_steps = max_steps
        self.reference_semantics = bool(reference_semantics)
        self.reference_temperature = bool(reference_temperature) or self.reference_semantics
        self.cost_fn = cost_fn      # replaces the simulations (tests: the trainer logic without a GPU)
        self.simulator = None if cost_fn is not None else Collision_Avoidance_Sim(
            numAgents=numAgents, scenario=scenario, visualize=False, num_envs=self.chains * self.sims_per_eval,
            seed=seed, device=device)
        # one working set + one best set per chain (Train_ALAN_action_space.py:16-20)
        self.actions: List[List[Action]] = []
        for c in range(self.chains):
            self._chain = c
            self.actions.append([(1, 0), self.random_action()])
        self.actions_opt = list(self.actions) if self.reference_semantics else [list(a) for a in self.actions]
        self.eval = self.evaluate_action(self.actions)
        self.eval_opt = list(self.eval)
        self.init_temp, self.final_temp = 0.9, 0.1
        self.temp = self.init_temp
        self.delta_temp = (self.final_temp - self.init_temp) / max(1, self.numRounds - 1)
        self.history = []

    # ------------------------------------------------------------------ search loop
    def train(self):
        for i in range(self.numRounds):
            proposals, dists = [], []
            for c in range(self.chains):
                self._chain = c
                modification = self.select_modification(self.actions[c], i)
                # the reference hands its working list to the move, which edits it in place (:33)
                work = self.actions[c] if self.reference_semantics else list(self.actions[c])
                d, new_actions = self.apply_modification(work, modification)
                proposals.append(new_actions)
                dists.append(d)
            new_eval = self.evaluate_action(proposals, i)
            for c in range(self.chains):
                self._chain = c
                if new_eval[c] < self.eval_opt[c]:
                    self.actions_opt[c] = proposals[c] if self.reference_semantics else list(proposals[c])
                    self.eval_opt[c] = new_eval[c]
                accept = self.symmetric_likelihood(dists[c]) * exp(min(50.0, (self.eval[c] - new_eval[c]) / self.temp))
                if self.py_rngs[c].uniform(0, 1) < accept:
                    self.actions[c], self.eval[c] = proposals[c], new_eval[c]
            self.history.append((i, self.temp, min(self.eval_opt)))
            self.temp = self.temp - self.delta_temp if self.reference_temperature else self.temp + self.delta_temp
        best = int(np.argmin(self.eval_opt))
        return self.actions_opt[best]

    # ------------------------------------------------------------------ pieces (same names as the reference)
    @property
    def _np(self) -> np.random.RandomState:
        return self.np_rngs[self._chain]

    def random_action(self) -> Action:
        angle = self.py_rngs[self._chain].uniform(-pi, pi)
        return cos(angle), sin(angle)

    def evaluate_action(self, actions: Sequence[Sequence[Action]], i: int = 0) -> List[float]:
        """Mean TTime over ``sims_per_eval`` simulations for every chain's candidate, one batch."""
        if self.cost_fn is not None:
            return [float(self.cost_fn(a)) for a in actions]
        per_world = [list(a) for a in actions for _ in range(self.sims_per_eval)]
        self.simulator.reset(online_actions=per_world)
        finished, total_time, ttime, min_ttime = self.simulator.run_sim(mode=1, max_steps=self.max_steps)
        t = ttime.reshape(self.chains, self.sims_per_eval).mean(1)
        return [float(x) for x in t.cpu()]

    def select_modification(self, actions, i) -> int:
        if len(actions) <= 1:
            return 2
        return int(self._np.choice(3, p=[0.8, 0.1, 0.1]))

    def apply_modification(self, actions, modification):
        """:78-84.  One guard the reference does not need: the kernel's action table holds at
        most ``MAX_ACTIONS`` entries, so an 'add' on a full set becomes an 'edit'."""
        from . import _lib
        if modification == 2 and len(actions) < _lib.MAX_ACTIONS:
            return self.mod_add(actions)
        if modification == 1:
            return self.mod_remove(actions)
        return self.mod_edit(actions)

    def mod_edit(self, actions):
        index = int(self._np.choice(range(1, len(actions))))
        angle = np.arctan2(actions[index][1], actions[index][0])
        new_angle = self._np.normal(angle, pi)
        new_action = (cos(new_angle), sin(new_angle))
        d = self.dist(actions[index], new_action)
        actions[index] = new_action
        return d, actions

    def mod_remove(self, actions):
        """:100-112.  ``list.remove(old_action)`` drops the FIRST equal entry, as in the reference."""
        index = int(self._np.choice(range(1, len(actions))))
        old = actions[index]
        actions.remove(old)
        return min([10] + [self.dist(a, old) for a in actions]), actions

    def mod_add(self, actions):
        index = int(self._np.choice(range(0, len(actions))))
        angle = np.arctan2(actions[index][1], actions[index][0])
        new_angle = self._np.normal(angle, pi)
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
