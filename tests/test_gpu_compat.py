"""Drop-in check at the rvo2 boundary: every call the UNMODIFIED reference shells
made to ``rvo2.PyRVOSimulator`` while they ran in the build container -- constructor, addAgent,
addObstacle, processObstacles, setAgentPrefVelocity, doStep, the position / velocity / neighbor /
obstacle-vertex getters, setAgentPosition -- was recorded together with what it returned
(tests/golden/shell_calltrace.json.gz, made by tests/golden/make_shell_golden.py with the CPU
oracle behind the boundary).  Here the same call sequence is issued against
``collision_avoidance_b200.rvo2_compat.PyRVOSimulator`` (CUDA) and every return value must
match: ids exactly, floats within the north star's 1e-4.  The reference tree itself is not
needed (and does not exist) on the GPU box.

Traces: ALAN ``online_step`` + ``done_test`` in the `blocks` world, ``orca_step`` in `congested`
(two-stage goals), and the gym env's ``step`` / ``reset`` / ``orca_step`` with its observation
getters (collision_avoidence_env.py:231-318)."""
import numpy as np
import pytest

from _golden import load_calltrace, replay_calltrace

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("trace", ["alan_blocks_online", "alan_congested_orca", "env_step_reset_orca"])
def test_reference_shell_call_trace_replays_on_the_cuda_shim(trace):
    from collision_avoidance_b200 import rvo2_compat
    calls = load_calltrace()[trace]
    n, worst = replay_calltrace(calls, rvo2_compat.PyRVOSimulator, atol=1e-4)
    kinds = {c[0] for c in calls}
    print(f"{trace}: {n} boundary calls ({len(kinds)} distinct methods), worst float error {worst:.3g}")
    assert {"addAgent", "addObstacle", "processObstacles", "doStep", "setAgentPrefVelocity", "getAgentPosition"} <= kinds
    # the kernels are FMA-free and the shim feeds them the reference's own float32 inputs: bit-equal
    assert worst == 0.0


def test_shim_as_the_rvo2_module():
    """``sys.modules['rvo2'] = rvo2_compat`` is all a reference checkout needs (INTEGRATION.md
    section 1): the module exposes PyRVOSimulator with the upstream constructor signature."""
    import inspect
    from collision_avoidance_b200 import rvo2_compat
    sig = inspect.signature(rvo2_compat.PyRVOSimulator.__init__)
    assert list(sig.parameters)[1:9] == ["timeStep", "neighborDist", "maxNeighbors", "timeHorizon", "timeHorizonObst",
                                         "radius", "maxSpeed", "velocity"]
    sim = rvo2_compat.PyRVOSimulator(1 / 60., 1.5, 5, 1.5, 2, 0.4, 2)          # the env's 7 positionals (Q1)
    a = sim.addAgent((0.0, 0.0))
    b = sim.addAgent((3.0, 0.0), 1.5, 5, 1.5, 2.0, 0.4, 2.0, (0.0, 0.0))      # same parameters, spelled out
    assert (a, b) == (0, 1)
    with pytest.raises(ValueError):
        sim.addAgent((1.0, 1.0), 1.5, 5)
    sim.setAgentPrefVelocity(a, (1.0, 0.0))
    sim.doStep()
    assert sim.getAgentVelocity(a) == (1.0, 0.0) and sim.getAgentPosition(a)[0] == pytest.approx(1 / 60.)
    assert sim.getAgentNumAgentNeighbors(a) == 0
