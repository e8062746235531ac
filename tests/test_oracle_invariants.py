"""Implementation-independent invariants of the oracle.  Since upstream rvo2 is not
available, these pin the semantics: the LP result is the true constrained optimum, neighbor
lists are the brute-force sorted k-nearest with a strict range test, and the step commutes with
mirror symmetry."""
import numpy as np
import pytest

from _common import goal_pref, oracle_sims
from collision_avoidance_b200 import scenarios
from oracle.rvo2_oracle import solve_lp


def _brute_lp(lines, radius, pref):
    """Exact optimum of min |v - pref| s.t. half-planes and |v| <= radius, by enumerating the
    candidate vertices (float64)."""
    L = lines.astype(np.float64)
    P, D = L[:, :2], L[:, 2:]

    def viol(v):
        return D[:, 0] * (P[:, 1] - v[1]) - D[:, 1] * (P[:, 0] - v[0])

    pref = np.asarray(pref, np.float64)
    p0 = pref if np.linalg.norm(pref) <= radius else pref / np.linalg.norm(pref) * radius
    cands = [p0]
    for i in range(len(L)):
        cands.append(P[i] + np.dot(D[i], pref - P[i]) * D[i])
        dp = np.dot(P[i], D[i])
        disc = dp * dp + radius * radius - np.dot(P[i], P[i])
        if disc >= 0:
            cands += [P[i] + (-dp + s * np.sqrt(disc)) * D[i] for s in (-1, 1)]
        for j in range(i):
            den = D[i, 0] * D[j, 1] - D[i, 1] * D[j, 0]
            if abs(den) > 1e-12:
                num = D[j, 0] * (P[i, 1] - P[j, 1]) - D[j, 1] * (P[i, 0] - P[j, 0])
                cands.append(P[i] + num / den * D[i])
    best = None
    for c in cands:
        if np.linalg.norm(c) <= radius + 1e-6 and (len(L) == 0 or (viol(c) <= 1e-6).all()):
            d = np.linalg.norm(c - pref)
            if best is None or d < best[0]:
                best = (d, c)
    return best


def test_lp_result_is_the_constrained_optimum():
    scn = scenarios.circle(3, 16, seed=0)
    sims = oracle_sims(scn)
    feasible = infeasible = 0
    for t in range(120):
        for e, s in enumerate(sims):
            pref = goal_pref(s.positions(), scn.goal[e]).astype(np.float32)
            s.set_pref_velocities(pref)
            s.doStep()
            if t % 6:
                continue
            for i in range(16):
                lines, n_obst = s.orca_lines(i)
                fail, res = solve_lp(lines, n_obst, 1.0, pref[i])
                assert np.array_equal(res, np.array(s.getAgentVelocity(i), np.float32))
                best = _brute_lp(lines, 1.0, pref[i])
                if fail == len(lines):
                    feasible += 1
                    assert best is not None
                    assert np.linalg.norm(best[1] - res) < 1e-4
                    # every emitted half-plane is satisfied by the result
                    v = res.astype(np.float64)
                    pen = lines[:, 2] * (lines[:, 1] - v[1]) - lines[:, 3] * (lines[:, 0] - v[0])
                    assert (pen <= 1e-4).all()
                else:
                    infeasible += 1
                    assert best is None  # LP2 reports infeasible only when the region is empty
                assert np.linalg.norm(res) <= 1.0 + 1e-4
    assert feasible > 100 and infeasible > 100


def test_neighbor_lists_are_brute_force_k_nearest():
    for N, k, nd in ((16, 10, 5.0), (60, 10, 5.0), (90, 5, 1.5)):
        scn = scenarios.crowd(1, N, seed=N)
        scn.params = dict(scn.params, maxNeighbors=k, neighborDist=nd)
        s = oracle_sims(scn)[0]
        for _ in range(10):
            pos = s.positions()
            s.set_pref_velocities(goal_pref(pos, scn.goal[0]).astype(np.float32))
            s.doStep()
            for i in range(N):
                d = (pos - pos[i]).astype(np.float32)
                d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)
                d2[i] = np.inf
                cand = np.where(d2 < np.float32(nd) ** 2)[0]
                expect = sorted(cand, key=lambda j: d2[j])[:k]
                got = s.agent_neighbors(i)
                assert [round(x[1], 9) for x in got] == [round(float(d2[j]), 9) for j in expect]
                assert sorted(x[0] for x in got) == sorted(expect) or len(set(d2[expect])) < len(expect)


def test_mirror_equivariance():
    """Reflecting the world in the x axis reflects the result (and reverses polygon winding)."""
    scn = scenarios.crowd(1, 30, seed=3, blocks=2)
    a = oracle_sims(scn)[0]
    m = scenarios.crowd(1, 30, seed=3, blocks=2)
    flip = np.array([1.0, -1.0], np.float32)
    m.pos, m.vel, m.goal = m.pos * flip, m.vel * flip, m.goal * flip
    m.obstacles = [[[(x, -y) for x, y in poly][::-1] for poly in m.obstacles[0]]]
    b = oracle_sims(m)[0]
    for _ in range(30):
        pa = goal_pref(a.positions(), scn.goal[0]).astype(np.float32)
        a.set_pref_velocities(pa)
        b.set_pref_velocities(pa * flip)
        a.doStep()
        b.doStep()
        # mirrored runs may differ in the last bits (left/right legs swap roles): tolerance, not equality
        assert np.abs(a.velocities() - b.velocities() * flip).max() < 1e-4
        b.set_positions(a.positions() * flip)
        b.set_velocities(a.velocities() * flip)


def test_feasible_pref_velocity_is_kept():
    scn = scenarios.crowd(1, 2, seed=1)
    scn.pos[0] = [[5.0, 5.0], [25.0, 25.0]]  # far apart, away from walls (envsize ~2.83 -> outside, no wall nearby)
    s = oracle_sims(scn)[0]
    s.set_pref_velocities(np.array([[0.3, -0.2], [0.0, 0.5]], np.float32))
    s.doStep()
    assert np.allclose(s.velocities(), [[0.3, -0.2], [0.0, 0.5]], atol=1e-7)
