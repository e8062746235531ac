"""Row f2: the shared policy network (run_rllib.py:35-52) on the CUDA path, against a float32
torch evaluation of the same layers; and the closed loop obs -> policy -> env.step on the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4  # float32 MLP, sums of 64 terms of O(1): summation order is the only difference


def _reference(policy, obs):
    import torch
    with torch.no_grad():
        h = torch.relu(obs.double() @ policy.w1.double() + policy.b1.double())
        h = torch.relu(h @ policy.w2.double() + policy.b2.double())
        return (h @ policy.w3.double() + policy.b3.double()).float()


def _sim():
    from collision_avoidance_b200.sim import BatchedRVOSimulator
    return BatchedRVOSimulator(2, 8, 1 / 60., 5.0, 10, 1.5, 1.5, 0.5, 1.0)


@pytest.mark.parametrize("rows", [1, 7, 127, 128, 129, 1000, 40_001])
@pytest.mark.parametrize("n_out", [1, 2, 8])
def test_policy_mlp_matches_torch(rows, n_out):
    import torch
    from collision_avoidance_b200.policy import SharedMLPPolicy
    pol = SharedMLPPolicy(_sim(), num_outputs=n_out, seed=rows + n_out)
    g = torch.Generator().manual_seed(rows)
    for b in (pol.b1, pol.b2, pol.b3):   # non-zero biases so they are exercised
        b.copy_(torch.randn(b.shape, generator=g) * 0.3)
    obs = (torch.randn(rows, 64, generator=g) * 2.0).cuda()
    y = pol(obs)
    ref = _reference(pol, obs)
    assert y.shape == (rows, n_out)
    assert float((y - ref).abs().max()) <= TOL
    # float32 torch evaluation for the record (same tolerance)
    h = torch.relu(obs @ pol.w1 + pol.b1)
    h = torch.relu(h @ pol.w2 + pol.b2)
    assert float((y - (h @ pol.w3 + pol.b3)).abs().max()) <= TOL


def test_policy_argument_errors():
    import torch
    from collision_avoidance_b200.policy import SharedMLPPolicy
    pol = SharedMLPPolicy(_sim())
    with pytest.raises(ValueError):
        pol(torch.zeros(4, 64))                      # CPU tensor
    with pytest.raises(ValueError):
        pol(torch.zeros(4, 63, device="cuda"))       # wrong width
    with pytest.raises(ValueError):
        pol.load_state({k: v[:1] for k, v in pol.state().items()})
    with pytest.raises(NotImplementedError):
        SharedMLPPolicy(_sim(), num_outputs=9)


def test_closed_loop_on_device():
    """obs -> policy -> env.step for a batch of gym worlds, no host round trip; the policy's
    actions are what the env consumed (same rewards as stepping with the tensor explicitly)."""
    import torch
    from collision_avoidance_b200 import envs
    from collision_avoidance_b200.policy import SharedMLPPolicy
    a = envs.Collision_Avoidance_Env(numAgents=10, num_envs=64, seed=11)
    b = envs.Collision_Avoidance_Env(numAgents=10, num_envs=64, seed=11)
    pol = SharedMLPPolicy(a.sim, seed=3)
    obs_a, obs_b = a.reset(), b.reset()
    for _ in range(30):
        theta = pol.act(obs_a)
        assert theta.shape == (64, 10)
        ref = _reference(pol, obs_b.reshape(-1, 64))[:, 0].reshape(64, 10).clamp(-np.pi, np.pi)
        assert float((theta - ref).abs().max()) <= TOL
        obs_a, rew_a, done_a, _ = a.step(theta)
        obs_b, rew_b, done_b, _ = b.step(theta.clone())
        assert torch.equal(obs_a, obs_b) and torch.equal(rew_a, rew_b)
