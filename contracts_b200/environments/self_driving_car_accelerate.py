"""`SelfAcceleratingCarEnv` with the reference's constructor and dict API
(environments/self_driving_car_accelerate.py:18-250), stepping on the GPU.

`collision_on=True` (never set by a shipped config) is rejected; randomness is the counter-based Philox stream
(kwargs `seed`, `env_id`).  Agents that are done are dropped from the returned dicts once their done flag has been
reported, and their actions are ignored — what RLlib's sampler does with the reference env.
"""
import numpy as np
import torch

from .. import spaces
from ..selfdrive import BatchedCarEnv

ACCEL_LOW_THRESH, ACCEL_HIGH_THRESH = -0.1, 0.1


class SelfAcceleratingCarEnv:
    def __init__(self, low_bound=-10.0, high_bound=10.0, start_vel=0.2, start_vel_ambulance=0.8, num_agents=2,
                 collision_on=False, num_envs=1, seed=73907, env_id=0, device=None, **kwargs):
        if collision_on:
            raise NotImplementedError("collision_on=True is not part of the accelerated path")
        self.num_agents = num_agents
        self.low_bound, self.high_bound = low_bound, high_bound
        self.start_vel, self.start_vel_ambulance = start_vel, start_vel_ambulance
        self.collision_on = collision_on
        self.num_envs, self.seed, self.env_id, self.device = int(num_envs), int(seed), int(env_id), device
        self.agent_ids = ["a%d" % i for i in range(num_agents)]
        self.agent_dones = {k: False for k in self.agent_ids}
        self.agent_dones["__all__"] = False
        self.observation_space = spaces.Box(low=self.low_bound - 20, high=self.high_bound + 20,
                                            shape=(2 * (self.num_agents + 1) + 3,), dtype=np.float32)
        self.action_space = spaces.Box(low=ACCEL_LOW_THRESH, high=ACCEL_HIGH_THRESH, shape=(1,), dtype=np.float32)
        self._contract = None
        self._batch = None
        self._last = None

    @property
    def batch(self):
        if self._batch is None:
            c = self._contract or (None, 0.0, 100.0, 0.0)
            self._batch = BatchedCarEnv(self.num_envs, self.num_agents, contract=c[0], low_bound=self.low_bound,
                                        high_bound=self.high_bound, start_vel=self.start_vel,
                                        start_vel_ambulance=self.start_vel_ambulance, theta_low=c[1], theta_high=c[2],
                                        null_prob=c[3], seed=self.seed, first_env_id=self.env_id, device=self.device)
        return self._batch

    def _bind_contract(self, name, low, high, null_prob):
        if self._batch is not None:
            self._batch.close()
            self._batch = None
        self._contract = (name, float(low), float(high), float(null_prob))

    @property
    def metrics(self):
        return {"transfers": float(self.batch.get_state()["transfers"][0].item())}

    @property
    def agent_positions(self):
        p = self.batch.get_state()["pos"][0].cpu().numpy()
        return {k: float(p[i]) for i, k in enumerate(self.agent_ids)}

    def reset(self):
        obs = self.batch.reset()[0].cpu().numpy()
        self.agent_dones = {k: False for k in self.agent_ids}
        self.agent_dones["__all__"] = False
        return {k: obs[i].copy() for i, k in enumerate(self.agent_ids)}

    def step(self, acts):
        if self.agent_dones["__all__"]:
            raise RuntimeError("episode is over: call reset() (the reference raises AttributeError here, :160)")
        b = self.batch
        # The acting set is the device's (cars that are done stop acting, as under RLlib): the dict must name exactly the
        # live cars.  The reference would move only the cars that are named and crash on a finished one; a mismatch here is
        # a caller error and is reported instead of being silently reinterpreted.
        live = {k for k in self.agent_ids if not self.agent_dones.get(k, False)}
        if set(acts.keys()) != live:
            raise ValueError("selfdrive step: actions for %s given, live cars are %s" % (sorted(acts.keys()), sorted(live)))
        # accelerations are float32 on the device — the reference's action space is a float32 Box, so RLlib hands float32
        # values over; a float64 action that float32 cannot represent is rounded before the +-0.1 clamp
        a = np.zeros((self.num_envs, self.num_agents), dtype=np.float32)
        for k, v in acts.items():
            a[0, int(k[1:])] = np.float32(np.asarray(v).reshape(-1)[0])
        obs, rew, done, info = b.step(torch.from_numpy(a).to(b.device))
        L = {"obs": obs[0].cpu().numpy(), "rew": rew[0].cpu().numpy(), "base_rew": b.base_rew[0].cpu().numpy(),
             "transfers": b.transfers[0].cpu().numpy(), "info": info[0].cpu().numpy(), "done": done[0].cpu().numpy()}
        self._last = L
        acting = [k for i, k in enumerate(self.agent_ids) if L["info"][i, 1] == 1.0]
        obs_d = {k: L["obs"][int(k[1:])].copy() for k in acting}
        rews = {k: float(L["base_rew"][int(k[1:])]) for k in acting}
        for i, k in enumerate(self.agent_ids):
            self.agent_dones[k] = bool(L["done"][i])
        self.agent_dones["__all__"] = bool(L["done"][-1])
        infos = {}
        for j, k in enumerate(acting):
            i = int(k[1:])
            infos[k] = {"just_passed": bool(L["info"][i, 0]), "is_crashed": 0,
                        "ambulance_rank": L["info"][i, 2] if j == 0 else 0.0,
                        "ambulance_dist_to_front": L["info"][i, 3] if j == 0 else 0.0}
        return obs_d, rews, self.agent_dones, infos

    def render(self, mode="rgb"):
        return True

    def close(self):
        if self._batch is not None:
            self._batch.close()
            self._batch = None
