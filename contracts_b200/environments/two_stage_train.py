"""Contract wrappers with the reference's names and dict API (environments/two_stage_train.py:17-358).

The reward redistribution of `SeparateContractEnv.step` (:62-121) is not evaluated here: it is fused into the
step kernel (selected by the contract's class name), this module only shapes the dict outputs.
"""
import copy

import numpy as np
import torch

from .. import spaces
from ..batched import negotiate

_FUSED = ("CleanupContract", "HarvestFeaturemodLocalContract", "SelfdriveContractDistprop")


class SeparateContractEnv:
    """two_stage_train.py:17-127."""
    metadata = {"render.modes": ["rgb_array"]}

    def __init__(self, base_env, contract, num_agents, convolutional, env_params=None, null_prob=0.0, **kwargs):
        self.num_agents = num_agents
        self.base_env = base_env
        self.contract = contract
        self.contract_low = self.contract.contract_space.low
        self.contract_high = self.contract.contract_space.high
        self.contract_state = {"a" + str(i): 0 for i in range(self.num_agents)}
        self.convolutional = convolutional
        self.null_prob = null_prob
        name = type(contract).__name__
        if name not in _FUSED:
            raise NotImplementedError("contract %s has no fused device transfer function" % name)
        base_env._bind_contract(name, self.contract_low[0], self.contract_high[0], null_prob)
        self.agent_ids = ["a%d" % i for i in range(num_agents)]
        if self.convolutional:
            contract_space = spaces.Box(low=np.concatenate((self.contract_low, np.array([0.0]))),
                                        high=np.concatenate((self.contract_high, np.array([3.0]))))
            obs_space = self.base_env.observation_space
            d = {"contract": contract_space}
            for k in ("features", "image"):
                if k in obs_space.keys():
                    d[k] = obs_space[k]
            self.observation_space = spaces.Dict(d)
        else:
            self.observation_space = spaces.Box(
                low=np.concatenate((self.base_env.observation_space.low, self.contract_low, np.array([0.0]))),
                high=np.concatenate((self.base_env.observation_space.high, self.contract_high, np.array([3.0]))))
        self.params = None

    # -- helpers ------------------------------------------------------------------------------------
    def _theta(self):
        return self.base_env.batch.get_state()["theta"][0:1].cpu().numpy().astype(np.float64)

    def _with_contract(self, obs, state_of):
        out = {}
        for k in self.agent_ids:
            if k not in obs:            # selfdrive: cars that are done no longer act / observe
                continue
            tail = np.array([state_of(k)])
            if self.convolutional:
                out[k] = obs[k]
                out[k].update({"contract": np.concatenate((self.params[k], tail))})
            else:
                out[k] = np.concatenate((obs[k], self.params[k], tail))
        return out

    def step(self, acts):
        raw_obs, _, dones, infos = self.base_env.step(acts)
        L = self.base_env._last
        self.obs = {k: raw_obs[k] for k in acts.keys()}
        rews = {k: np.float64(L["rew"][int(k[1:])]) for k in acts.keys()}     # after transfers (kernel)
        self.last_transfers = {k: L["transfers"][int(k[1:])] for k in acts.keys()}
        for k in acts.keys():
            infos[k]["contract_param"] = self.params[k]
        return self._with_contract(self.obs, lambda k: 0), rews, dones, infos

    def render(self, mode="rgb"):
        return self.base_env.render()

    def reset(self):
        raise NotImplementedError

    @property
    def metrics(self):
        return self.base_env.metrics


class SeparateContractSubgameStage(SeparateContractEnv):
    """two_stage_train.py:129-187: every reset samples theta ~ U(low, high) (or `low` with prob. null_prob)."""

    def __init__(self, base_env, contract, num_agents, convolutional, env_params=None, null_prob=0.0, **kwargs):
        super().__init__(base_env, contract, num_agents, convolutional, env_params, null_prob)
        self.action_space = self.base_env.action_space

    def reset(self):
        base_obs = self.base_env.reset()           # the reset kernel also draws theta (two_stage_train.py:163-166)
        self.obs = copy.deepcopy(base_obs)
        rand_val = self._theta()
        self.contract_state = {k: 0 for k in self.agent_ids}
        self.params = {k: rand_val for k in self.agent_ids}
        return self._with_contract(self.obs, lambda k: 0)


class SeparateContractNegotiateStage(SeparateContractEnv):
    """two_stage_train.py:190-358: propose (state 2) -> accept / reject (state 3) -> frozen-policy rollout.

    The reference loads a frozen RLlib PPO policy (`trainer_config`, `trainer_env`, `trainer_path`); that
    forward pass is outside the accelerated path.  Here `policy` is any callable
    `policy(obs, agent_id) -> action` playing that role; with `policy=None` the subgame is rolled out with
    uniform random actions generated on the device (what the throughput benchmark uses).
    """

    def __init__(self, base_env, contract, num_agents, horizon, trainer_config=None, trainer_env=None,
                 trainer_path=None, convolutional=True, shared=True, env_params=None, policy=None, **kwargs):
        super().__init__(base_env, contract, num_agents, convolutional)
        self.horizon = horizon
        self.shared = shared
        self.policy = policy
        self.action_space = spaces.Box(low=np.concatenate((self.contract_low, np.array([0.0]))),
                                       high=np.concatenate((self.contract_high, np.array([1.0]))))
        self._metrics = {"contract": -1, "accepted": 0}

    @property
    def metrics(self):
        return self._metrics

    def reset(self):
        self._metrics = {"contract": -1, "accepted": 0}
        base_obs = self.base_env.reset()
        self.obs = copy.deepcopy(base_obs)
        self.last_seen_obs = copy.deepcopy(base_obs)
        self.params = None
        self.contract_state = {k: 2 for k in self.agent_ids}
        zeros = {k: np.zeros(self.contract_low.shape) for k in self.agent_ids}
        saved, self.params = self.params, zeros
        out = self._with_contract(self.obs, lambda k: self.contract_state[k])
        self.params = saved
        return out

    def step(self, acts):
        b = self.base_env.batch
        n = self.num_agents
        if self.contract_state["a0"] == 2:
            proposal = np.asarray(acts["a0"][:-1], dtype=np.float64)
            self._metrics["contract"] = acts["a0"][:-1]
            self.params = {k: proposal for k in acts.keys()}
            self.contract_state = {k: 3 for k in self.agent_ids}
            rews = {k: 0.0 for k in self.agent_ids}
            dones = {"__all__": False}
            infos = {k: {} for k in self.agent_ids}
        elif self.contract_state["a0"] == 3:
            accept = torch.tensor([[float(acts[k][-1]) for k in self.agent_ids]], dtype=torch.float64)
            dec = negotiate(b, torch.tensor([float(self.params["a0"][0])], dtype=torch.float64), accept)
            decision = int(dec[0].item())
            self._metrics["accepted"] = decision
            for k in self.agent_ids:
                self.contract_state[k] = 0
                self.params[k] = self.params[k] if decision == 1 else np.zeros(shape=self.contract_low.shape)
            rews = {k: 0.0 for k in self.agent_ids}
            dones = {"__all__": True}
            infos = {k: {} for k in self.agent_ids}
            if self.policy is None:
                from .gridworld import _GridWorldEnv
                if not isinstance(self.base_env, _GridWorldEnv):       # the feature envs also carry `image_obs`
                    raise NotImplementedError("the random-action subgame rollout exists for the grid worlds; pass policy=...")
                rews, infos = self._random_rollout(rews, infos)
            else:
                env_done, steps = False, 0
                active = list(self.agent_ids)            # agents that finished stop acting (selfdrive; :286,325-333)
                while not env_done and steps < self.horizon:
                    act_dict = {}
                    for k in active:
                        if self.convolutional:
                            o = self.obs[k]
                            o.update({"contract": np.concatenate((self.params[k], np.array([0])))})
                        else:
                            o = np.concatenate((self.obs[k], self.params[k], np.array([0])))
                        act_dict[k] = self.policy(o, k)
                    _, env_rews, env_dones, infos = super().step(act_dict)
                    self.last_seen_obs = {k: self.obs[k] if k in self.obs else self.last_seen_obs[k] for k in self.agent_ids}
                    env_done = env_dones["__all__"]
                    steps += 1
                    for k in active:
                        rews[k] += env_rews[k]
                    active = [k for k in active if not env_dones.get(k, False)]
        else:
            raise RuntimeError("negotiation episode is over: call reset()")
        saved_obs = self.obs
        out = self._with_contract(self.last_seen_obs, lambda k: self.contract_state[k])
        self.obs = saved_obs
        return out, rews, dones, infos

    def _random_rollout(self, rews, infos):
        """<= horizon subgame steps with device-generated uniform random actions; rewards summed on the device."""
        b = self.base_env.batch
        na = self.base_env.action_space.n
        total = torch.zeros_like(b.rew)
        # the reference stops at env_dones['__all__'] (:286-333), i.e. when the base env reaches ITS horizon: the number
        # of steps is known up front, so the loop needs no device -> host poll of the done flag
        limit = min(int(self.horizon), max(int(self.base_env.horizon) - int(self.base_env.timesteps), 0))
        for steps in range(limit):
            a = b.random_actions(steps, na)
            b.step(a, want_features=False)
            total += b.rew
            self.base_env.timesteps += 1
        tot = total[0].cpu().numpy()
        rews = {k: rews[k] + tot[i] for i, k in enumerate(self.agent_ids)}
        self.last_seen_obs = self.base_env._image_obs(b.obs[0].cpu().numpy()) if self.base_env.image_obs else self.last_seen_obs
        for k in self.agent_ids:
            infos[k]["contract_param"] = self.params[k]
        return rews, infos


class JointEnv:
    """two_stage_train.py:476-617: one centralised controller 'a0' acts for every agent.

    `global_obs`: the observation is the whole colour map (`base_env.get_global_obs()`); `concatenated_obs`: the
    agents' windows concatenated along the channel axis; otherwise (`duplicate_obs` or not) the flat-observation
    variant for Box action spaces (selfdrive).  The image layouts are written on the device
    (`ssd_global_view` / `ssd_concat_obs`); rewards are summed, infos summed key by key, as in the reference.
    """

    def __init__(self, base_env, num_agents=2, duplicate_obs=False, concatenated_obs=False, global_obs=False, **kwargs):
        self.base_env, self.num_agents = base_env, num_agents
        self.duplicate_obs, self.concatenated_obs, self.global_obs = duplicate_obs, concatenated_obs, global_obs
        self._ids = ["a%d" % i for i in range(num_agents)]
        if global_obs or concatenated_obs:                       # image layouts: MultiDiscrete joint action (:503-508)
            self.observation_space = (base_env.global_observation_space if global_obs
                                      else base_env.concatenated_observation_space)
            self.action_space = base_env.global_action_space
        else:                                                    # flat layouts: tiled Box spaces (:509-523)
            tile = lambda x: np.concatenate([x] * num_agents)    # noqa: E731
            obs_space = base_env.observation_space
            self.observation_space = spaces.Box(low=tile(obs_space.low), high=tile(obs_space.high)) if duplicate_obs else obs_space
            self.action_space = spaces.Box(low=tile(base_env.action_space.low), high=tile(base_env.action_space.high))
            self.curr_agent_lst = list(self._ids)

    @property
    def metrics(self):
        return self.base_env.metrics

    def _image_obs(self):
        if self.global_obs:
            return {"a0": self.base_env.get_global_obs()}
        img = self.base_env.batch.concatenated_obs()[0].cpu().numpy()
        return {"a0": {"image": img.astype(np.float64) / 255}}

    @staticmethod
    def _joint(obs, env_rews, env_dones, env_infos, members):
        """Summed reward (not the average, :592), the env-wide done flag, infos summed key by key (:594-595)."""
        infos = {key: sum([env_infos[k][key] for k in members]) for key in env_infos[members[0]].keys()}
        return obs, {"a0": sum([r for r in env_rews.values()])}, {"a0": env_dones["__all__"], "__all__": env_dones["__all__"]}, {"a0": infos}

    def reset(self):
        first = self.base_env.reset()
        if self.global_obs or self.concatenated_obs:
            return self._image_obs()
        self.agent_obs = dict(first)                             # assumes every agent is in the initial observation
        self.curr_agent_lst = list(self._ids)
        return {"a0": np.concatenate([first[k] for k in self._ids])}

    def step(self, acts):
        joint = acts["a0"]
        if self.global_obs or self.concatenated_obs:             # global_step / concatenated_step (:572-617)
            _, rews, dones, infos = self.base_env.step({k: joint[i] for i, k in enumerate(self._ids)})
            return self._joint(self._image_obs(), rews, dones, infos, self._ids)
        width = self.base_env.action_space.shape[0]              # each active agent's slice of the joint action (:540-546)
        active = list(self.curr_agent_lst)
        base_obs, rews, dones, infos = self.base_env.step(
            {k: np.array(joint[i * width:(i + 1) * width]) for i, k in enumerate(self._ids) if k in active})
        for k in active:
            self.agent_obs[k] = base_obs[k]
        obs = np.concatenate([self.agent_obs[k] for k in self._ids]) if self.duplicate_obs else self.agent_obs[active[0]]
        out = self._joint({"a0": obs}, rews, dones, infos, active)
        self.curr_agent_lst = [k for k in active if not dones.get(k, False)]     # finished agents stop acting (:562-566)
        return out

    def render(self, mode="rgb"):
        return self.base_env.render()


class NegotiationSolver(SeparateContractEnv):
    """two_stage_train.py:619-776: at every reset, sample `contract_samples` contracts, query the agents' frozen value
    functions for each (plus the null contract) and keep the best one under `decision_rule` ('max' | 'majority').

    The reference loads a frozen RLlib PPO policy (`trainer_config`, `trainer_env`, `trainer_path`) and reads
    `model.value_function()` after `compute_single_action` (:693-703); that forward pass is outside the accelerated
    path.  Here `value_fn(obs, agent_id) -> float` plays that role (obs = the agent's observation with the candidate
    contract appended, exactly what the reference feeds its policy).  Candidate sampling and the decision rule run on
    the device (`ssd_solver_sample` / `ssd_solver_choose`).  The chosen parameter is float32-valued like gym's
    `Box.sample()`; the transfers are float64 arithmetic on it, as under the NumPy the reference pins.
    """

    def __init__(self, base_env, contract, num_agents, horizon=1000, trainer_config=None, trainer_env=None, trainer_path=None,
                 convolutional=True, shared=True, env_params=None, contract_samples=50, decision_rule="majority",
                 value_fn=None, **kwargs):
        super().__init__(base_env, contract, num_agents, convolutional)
        if value_fn is None:
            raise ValueError("NegotiationSolver needs value_fn(obs, agent_id) -> float (the frozen policy's value function)")
        if decision_rule not in ("max", "majority"):
            raise ValueError("decision_rule must be 'max' or 'majority'")
        self.horizon = horizon
        self.shared = shared
        self.value_fn = value_fn
        self.contract_param_space = spaces.Box(low=contract.contract_space.low, high=contract.contract_space.high)
        self.num_samples = contract_samples
        self.decision_rule = decision_rule
        self.config = trainer_config
        self.action_space = self.base_env.action_space
        self._metrics = {"contract": -1, "accepted": 0}

    @property
    def metrics(self):
        return self._metrics

    def _obs_with(self, param):
        out = {}
        for k in self.agent_ids:
            if self.convolutional:
                o = dict(self.obs[k])
                o["contract"] = np.concatenate((param, np.array([0])))
            else:
                o = np.concatenate((self.obs[k], param, np.array([0])))
            out[k] = o
        return out

    def compute_vals(self, obs):
        return {k: self.value_fn(obs[k], k) for k in obs}

    def negotiate(self):
        from ..batched import solver_choose, solver_sample
        b = self.base_env.batch                                          # any Batched*Env (grid, feature, selfdrive)
        params = solver_sample(b, self.num_samples)                      # [E, 1 + S] on the device
        cand = params[0].cpu().numpy()
        vals = np.zeros((1, self.num_samples + 1, self.num_agents))
        for c, theta in enumerate(cand):                                  # null contract first (:709-724)
            v = self.compute_vals(self._obs_with(np.array([theta])))
            vals[0, c] = [v[k] for k in self.agent_ids]
        full = torch.zeros((b.E, self.num_samples + 1, self.num_agents), dtype=torch.float64, device=b.device)
        full[0] = torch.as_tensor(vals[0])
        best, _ = solver_choose(b, params, full, self.decision_rule)
        return np.array([best[0].item()], dtype=np.float32)

    def reset(self):
        self._metrics = {"contract": -1, "accepted": 0}
        base_obs = self.base_env.reset()
        self.obs = copy.deepcopy(base_obs)
        self.last_seen_obs = copy.deepcopy(base_obs)
        self.params = None
        self.contract_state = {k: 0 for k in self.agent_ids}
        self.contract_param = self.negotiate()
        self.params = {k: self.contract_param for k in self.agent_ids}
        return self._with_contract(self.obs, lambda k: 0)
