"""CPU: the HOST logic of the drop-in classes — dict shaping, info / done keys, reward types, the contract observation, the
metrics, the negotiation state machine, flat observations, seeding — with the C oracle standing in for the device batch.

`OracleBatch` implements the slice of the `BatchedGridEnv` interface the drop-in classes call (reset / step / host_snapshot /
get_state / metrics_raw / render / global_view / concatenated_obs, CPU torch tensors) on top of `oracle.GridOracle`, and is
patched in for the device class; the SAME test bodies that run on the GPU (tests/test_dropin_api.py,
tests/test_negotiate_golden.py) then replay the reference's golden episodes through `env_creator` and the wrappers.  What this
proves is the Python layer, not the kernels: the device results themselves are compared with the oracle by the `-m gpu` suite.
(The oracle is test infrastructure: nothing under contracts_b200/ imports it.)"""
import numpy as np
import pytest
import torch

import golden_util as gu


class OracleBatch:
    def __init__(self, kind, num_envs, num_agents, ascii_map=None, horizon=1000, contract=None, theta_low=0.0, theta_high=None,
                 null_prob=0.0, seed=73907, first_env_id=0, device=None, padded_obs=False, **shaping):
        from oracle import oracle
        from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP
        assert kind in ("cleanup_new", "harvest_new")
        self.kind, self.E, self.n = kind, int(num_envs), int(num_agents)
        self.ascii_map = list(ascii_map) if ascii_map is not None else (CLEANUP_MAP if kind == "cleanup_new" else HARVEST_MAP)
        self.device = torch.device("cpu")
        self.o = oracle.GridOracle(kind[:-4], self.E, self.n, self.ascii_map, horizon=horizon, contract=contract,
                                   theta_low=theta_low, theta_high=theta_high, null_prob=null_prob, seed=seed,
                                   first_env_id=first_env_id, **shaping)
        self._oracle = oracle
        self.seed, self.first_env_id = int(seed), int(first_env_id)
        default_high = {"CleanupContract": 0.2, "HarvestFeaturemodLocalContract": 10.0}.get(contract, 0.0)
        self.theta_low = float(np.float32(theta_low))                      # gym Box bounds are float32 (contract_list.py:20,43)
        self.theta_high = float(np.float32(default_high if theta_high is None else theta_high))
        self.obs = None
        self.rew = torch.zeros((self.E, self.n), dtype=torch.float64)      # persistent, rewritten by every step (as on the device)
        self._last = None

    def reset(self, mask=None):
        self.obs = torch.from_numpy(self.o.reset(None if mask is None else mask.numpy()))
        return self.obs

    def step(self, actions, want_features=False, extras=True, **kw):
        r = self.o.step(actions.numpy().astype(np.int32), want_features=True)
        self._last = r
        self.obs = torch.from_numpy(r["obs"])
        self.rew.copy_(torch.from_numpy(r["rew"]))
        self.done = torch.from_numpy(r["done"])
        return self.obs, self.rew, self.done, torch.from_numpy(r["info"].astype(np.uint8))

    def host_snapshot(self, index=None, features=False, extras=True, obs=True):
        r = self._last
        out = {"rew": r["rew"], "info": r["info"].astype(np.uint8), "done": r["done"].astype(np.uint8), "obs": r["obs"],
               "base_rew": r["base_rew"], "transfers": r["transfers"], "feat": r["feature_obs"]}
        return {k: (v[index] if index is not None else v) for k, v in out.items()}

    def random_actions(self, step_index, num_actions, out=None):
        rng = np.random.RandomState(int(step_index or 0))
        return torch.from_numpy(rng.randint(0, num_actions, size=(self.E, self.n)).astype(np.uint8))

    def get_state(self):
        return {k: torch.from_numpy(np.asarray(v)) for k, v in self.o.get_state().items()}

    def metrics_raw(self):
        return torch.from_numpy(self.o.metrics_raw())

    def render(self):
        return torch.from_numpy(self.o.render())

    def global_view(self):
        return torch.from_numpy(self.o.global_view())

    def concatenated_obs(self):
        return torch.from_numpy(self._oracle.concatenated_obs(self.obs.numpy()))

    def record_beams(self, on=True):
        pass

    def close(self):
        pass


def _negotiate(batch, proposals, accept, mask=None, out=None):
    return torch.from_numpy(batch.o.negotiate(np.asarray(proposals, dtype=np.float64), np.asarray(accept, dtype=np.float64)))


def _solver_sample(batch, num_samples):
    rows = [batch._oracle.solver_candidates(batch.seed, batch.first_env_id + e, int(batch.o.episode[e]), batch.theta_low,
                                            batch.theta_high, int(num_samples)) for e in range(batch.E)]
    return torch.from_numpy(np.stack(rows))


def _solver_choose(batch, params, vals, rule="majority"):
    picks = [batch._oracle.solver_choose(params[e].numpy(), vals[e].numpy(), rule) for e in range(batch.E)]
    best = np.array([p[0] for p in picks], dtype=np.float64)
    batch.o.set_theta(best)
    return torch.from_numpy(best), torch.tensor([int(p[1]) for p in picks], dtype=torch.int32)


@pytest.fixture
def oracle_device(monkeypatch, oracle_lib):
    from contracts_b200 import batched
    from contracts_b200.environments import gridworld, two_stage_train
    monkeypatch.setattr(gridworld, "BatchedGridEnv", OracleBatch)
    monkeypatch.setattr(two_stage_train, "negotiate", _negotiate)
    monkeypatch.setattr(batched, "solver_sample", _solver_sample)          # imported by NegotiationSolver.negotiate at call time
    monkeypatch.setattr(batched, "solver_choose", _solver_choose)


# a sample of every fixture family of the dict API (all of them run on the GPU)
DICT_FIXTURES = [n for n in gu.fixture_names() if n in (
    "cleanup_n2", "cleanup_n8", "harvest_n4", "harvest_n8", "cleanup_cramped_n8", "harvest_cramped_n8", "cleanup_n4_collective",
    "harvest_n5_inequity", "cleanup_n8_deep", "cleanup_n8_nocontract", "cleanup_n5_short_horizon", "harvest_n8_short_horizon")] or gu.fixture_names()[:6]


@pytest.mark.parametrize("name", DICT_FIXTURES)
def test_dict_api_host_logic_replays_reference(oracle_device, name):
    import test_dropin_api as T
    T.test_dict_api_replays_reference(name)


@pytest.mark.parametrize("name", gu.fixture_names("flatobs_"))
def test_flat_observation_host_logic(oracle_device, name):
    import test_dropin_api as T
    T.test_flat_feature_observations(name)


@pytest.mark.parametrize("name", gu.fixture_names("negotiate_"))
def test_negotiate_stage_host_logic(oracle_device, name):
    import test_negotiate_golden as T
    T.test_dropin_negotiate_stage_matches_reference(name)


@pytest.mark.parametrize("name", gu.fixture_names("joint_"))
def test_joint_env_host_logic(oracle_device, name):
    import test_cuda_solver_joint as T
    T.test_joint_dict_api_replays_reference(name)


@pytest.mark.parametrize("name", gu.fixture_names("solver_"))
def test_negotiation_solver_host_logic(oracle_device, name):
    import test_cuda_solver_joint as T
    T.test_solver_dict_api_replays_reference(name)


def test_render_seed_and_agent_pos_host_logic(oracle_device):
    import test_dropin_api as T
    T.test_render_and_global_obs()
    T.test_seed_and_agent_pos()
    T.test_env_creator_rejects_unknown_and_out_of_scope()


def test_random_rollout_negotiate_stage(oracle_device):
    """policy=None: the subgame is rolled out with batch-generated random actions for min(horizon, base horizon) steps and the
    rewards are summed per agent (two_stage_train.py:283-333 with the frozen policy replaced)."""
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    n = 3
    base = env_creator("CleanupNew", dict(num_agents=n, env_params={}, image_obs=True, horizon=12, seed=7, env_id=3))
    env = env_creator("ContractWrapperNegotiate", dict(base_env=base, contract=contract_list.CleanupContract(n), num_agents=n,
                                                       horizon=50, convolutional=True, shared=True, policy=None))
    obs = env.reset()
    assert sorted(obs) == ["a0", "a1", "a2"] and all(o["contract"].tolist() == [0.0, 2.0] for o in obs.values())
    acts = {k: np.array([0.1, 1.0]) for k in obs}
    obs, rew, done, info = env.step(acts)
    assert done == {"__all__": False} and all(o["contract"].tolist() == [0.1, 3.0] for o in obs.values())
    obs, rew, done, info = env.step(acts)
    assert done == {"__all__": True} and env.metrics["accepted"] == 1
    assert base.timesteps == 12                                  # stopped at the base env's horizon, not at 50
    assert all(np.asarray(info[k]["contract_param"]).tolist() == [0.1] for k in obs)
    with pytest.raises(RuntimeError):
        env.step(acts)


# ---- feature envs and selfdrive: the same idea over FeatOracle / CarOracle ------------------------------------------------
class _FlatOracleBatch:
    """step() -> (obs, rew, done, info) CPU tensors + persistent base_rew / transfers, like BatchedFeatureEnv / BatchedCarEnv"""

    def _finish_init(self, action_dtype):
        self.device = torch.device("cpu")
        self._action_dtype = action_dtype
        self.base_rew = torch.zeros((self.E, self.n), dtype=torch.float64)
        self.transfers = torch.zeros((self.E, self.n), dtype=torch.float64)

    def reset(self, mask=None):
        return torch.from_numpy(self.o.reset(None if mask is None else mask.numpy()))

    def step(self, actions, extras=True, auto_reset=False):
        r = self.o.step(actions.numpy().astype(self._action_dtype))
        self.base_rew.copy_(torch.from_numpy(r["base_rew"]))
        self.transfers.copy_(torch.from_numpy(r["transfers"]))
        return torch.from_numpy(r["obs"]), torch.from_numpy(r["rew"]), torch.from_numpy(r["done"]), torch.from_numpy(np.asarray(r["info"]))

    def get_state(self):
        return {k: torch.from_numpy(np.asarray(v)) for k, v in self.o.get_state().items()}

    def metrics_raw(self):
        return torch.from_numpy(self.o.metrics_raw())

    def close(self):
        pass


class OracleFeatBatch(_FlatOracleBatch):
    def __init__(self, kind, num_envs, num_agents, ascii_map=None, horizon=1000, contract=None, theta_low=0.0, theta_high=None,
                 null_prob=0.0, seed=73907, first_env_id=0, device=None):
        from oracle import oracle
        self.E, self.n = int(num_envs), int(num_agents)
        self.o = oracle.FeatOracle(kind, self.E, self.n, list(ascii_map), horizon=horizon, contract=contract, theta_low=theta_low,
                                   theta_high=theta_high, null_prob=null_prob, seed=seed, first_env_id=first_env_id)
        self._finish_init(np.int32)


class OracleCarBatch(_FlatOracleBatch):
    def __init__(self, num_envs, num_agents, contract=None, low_bound=-10.0, high_bound=10.0, start_vel=0.2, start_vel_ambulance=0.8,
                 theta_low=0.0, theta_high=100.0, null_prob=0.0, seed=73907, first_env_id=0, device=None):
        from oracle import oracle
        self.E, self.n = int(num_envs), int(num_agents)
        self.o = oracle.CarOracle(self.E, self.n, contract=bool(contract), low_bound=low_bound, high_bound=high_bound,
                                  start_vel=start_vel, start_vel_ambulance=start_vel_ambulance, theta_low=theta_low,
                                  theta_high=theta_high, null_prob=null_prob, seed=seed, first_env_id=first_env_id)
        self._finish_init(np.float32)


@pytest.fixture
def oracle_flat_devices(monkeypatch, oracle_lib):
    from contracts_b200.environments import feature_envs, self_driving_car_accelerate
    monkeypatch.setattr(feature_envs, "BatchedFeatureEnv", OracleFeatBatch)
    monkeypatch.setattr(self_driving_car_accelerate, "BatchedCarEnv", OracleCarBatch)


@pytest.mark.parametrize("name", [n for n in gu.fixture_names("features_") if "nocontract" not in n])
def test_feature_env_dict_api_host_logic(oracle_flat_devices, name):
    import test_features_golden as T
    T.test_dropin_feature_env_dict_api(name)


@pytest.mark.parametrize("name", [n for n in gu.fixture_names("selfdrive_") if "nocontract" not in n and n != "selfdrive_n1"])
def test_selfdrive_dict_api_host_logic(oracle_flat_devices, name):
    import test_selfdrive_golden as T
    T.test_dropin_selfdrive_dict_api(name)


def test_selfdrive_dict_api_rejects_wrong_acting_set(oracle_flat_devices):
    """the dict must name exactly the live cars (the reference would move only the named ones / crash on a finished one)"""
    from contracts_b200.utils.env_creator_functions import env_creator
    env = env_creator("SelfDrive", dict(num_agents=3, seed=1, env_id=2))
    obs = env.reset()
    assert sorted(obs) == ["a0", "a1", "a2"]
    with pytest.raises(ValueError):
        env.step({"a0": [0.05], "a1": [0.05]})
    obs, rew, done, info = env.step({k: [0.05] for k in obs})
    assert set(done) == {"a0", "a1", "a2", "__all__"} and set(info["a0"]) == {"just_passed", "is_crashed", "ambulance_rank",
                                                                               "ambulance_dist_to_front"}
