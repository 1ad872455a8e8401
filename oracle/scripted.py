"""Scripted stand-ins for the frozen RLlib policy of the reference's NegotiationSolver.

TEST INFRASTRUCTURE (oracle/).  The reference queries `frozen_trainer.compute_single_action(obs)` followed by
`get_policy(id).model.value_function().item()` (two_stage_train.py:693-703); RLlib is not part of this image and
is outside the accelerated path, so the golden fixtures and the tests use this deterministic value function of
(observation, contract parameter, agent) instead.  It is shared by oracle/make_golden.py (driving the unmodified
reference) and by tests/ (driving the CUDA path), so both sides see identical values.
"""
import numpy as np


def scripted_value(obs, agent_index, num_agents, scale):
    """V_i(s, c): a concave function of theta (scaled to [0, 1]) whose optimum differs per agent, plus a
    state-dependent offset.

    obs: the convolutional observation dict {'image': float64 [15,15,3], 'contract': [theta, flag]}.
    Returns np.float64 (NOT a Python float: `sum()` of Python floats is compensated from Python 3.12 on,
    whereas the reference pins Python 3.9 — np.float64 keeps the welfare sums plain left-to-right adds)."""
    theta = np.float64(obs["contract"][0]) / np.float64(scale)
    img = np.asarray(obs["image"], dtype=np.float64)
    # agent 0 gains a lot from a large theta, the others mildly prefer a small one: the welfare maximum ('max' rule)
    # is a contract most agents like less than the null contract, so the 'majority' rule picks a different one
    peak = np.float64(0.9 if agent_index == 0 else 0.05 + 0.1 * (agent_index % 3))
    weight = np.float64(2.0 * num_agents if agent_index == 0 else 1.0 + 0.25 * agent_index)
    offset = np.float64(img[7, 7].sum() + 0.25 * img[:, :, 1].mean())
    return np.float64(offset + weight * (1.0 - (theta - peak) * (theta - peak)))
