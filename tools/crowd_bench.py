"""Wall-only crowds of 50 / 64 / 256 agents per world: slim (K + 2 line slots, 4 blocks per SM) vs standard tile kernel
(ORCA_B200_NO_SLIM_KERNEL=1).  Dev tool."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tools.bench_configs import orca_policy, timed
from collision_avoidance_b200 import scenarios
for E, N in ((20480, 50), (4096, 256), (16384, 64)):
    sim, step = orca_policy(scenarios.crowd(E, N, seed=2))
    print(E, N, "%.1f us" % (timed(step, 150, 30) * 1e3), sim.read_stats()["overflow"])
    del sim, step
