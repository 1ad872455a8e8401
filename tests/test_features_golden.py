"""CleanupFeatures / HarvestFeatures ('Cleanup' / 'Harvest' tags; cleanup_features.py:48-336,
harvest_features.py:60-364) + contract wrapper: the reference's golden episodes
(tests/golden/features_*.npz) replayed bit-exactly through the C oracle (CPU) and the CUDA path (GPU)."""
import numpy as np
import pytest

import golden_util as gu

NAMES = gu.fixture_names("features_")


def test_features_fixtures_present():
    assert len(NAMES) >= 5


def _maps(kind):
    from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP
    return CLEANUP_MAP if kind == "cleanup" else HARVEST_MAP


def _metrics(kind, n, raw):
    m = {"raw_env_rewards": raw[1], "transfers": raw[2]}
    if kind == "cleanup":
        m["dirt_cleaned"] = raw[0]
    else:
        m["total_apples_eaten"] = raw[3]
        m["low_density_apples_eaten"] = raw[4]
    return m


class _Oracle:
    def __init__(self, mod, fx):
        kind = str(fx["kind"])
        self.o = mod.FeatOracle(kind, 1, int(fx["n"]), _maps(kind), horizon=int(fx["horizon"]),
                                contract=gu.contract_name(fx), seed=int(fx["seed"]), first_env_id=int(fx["env_id"]))

    def reset(self):
        return self.o.reset()[0]

    def step(self, a):
        return {k: v[0] for k, v in self.o.step(np.asarray(a)[None]).items()}

    def state(self):
        return {k: v[0] for k, v in self.o.get_state().items()}

    def metrics_raw(self):
        return self.o.metrics_raw()[0]


class _Cuda:
    def __init__(self, fx, E=4, index=2):
        from contracts_b200.features import BatchedFeatureEnv
        kind = str(fx["kind"])
        self.i = index
        self.env = BatchedFeatureEnv(kind, E, int(fx["n"]), horizon=int(fx["horizon"]), contract=gu.contract_name(fx),
                                     seed=int(fx["seed"]), first_env_id=(int(fx["env_id"]) - index) & 0xFFFFFFFF)

    def reset(self):
        return self.env.reset()[self.i].cpu().numpy()

    def step(self, a):
        import torch
        acts = torch.zeros((self.env.E, self.env.n), dtype=torch.uint8)
        acts[:] = torch.as_tensor(np.asarray(a).astype(np.uint8))
        obs, rew, done, info = self.env.step(acts.cuda())
        i = self.i
        return {"obs": obs[i].cpu().numpy(), "rew": rew[i].cpu().numpy(), "base_rew": self.env.base_rew[i].cpu().numpy(),
                "transfers": self.env.transfers[i].cpu().numpy(), "info": info[i].cpu().numpy(), "done": int(done[i].item())}

    def state(self):
        return {k: v[self.i].cpu().numpy() for k, v in self.env.get_state().items()}

    def metrics_raw(self):
        return self.env.metrics_raw()[self.i].cpu().numpy()


def replay(backend, fx):
    kind, n = str(fx["kind"]), int(fx["n"])
    wrapped = bool(fx["contract"])
    F = 12 + n if kind == "cleanup" else 10 + 2 * n
    for ep in range(fx["actions"].shape[0]):
        ctx = "reset ep %d" % ep
        obs = backend.reset()
        st = backend.state()
        gu.assert_same("reset obs", obs, fx["reset_obs"][ep][:, :F], ctx)
        gu.assert_same("reset pos", st["pos"], fx["reset_pos"][ep], ctx)
        gu.assert_same("reset ori", st["ori"], fx["reset_ori"][ep], ctx)
        gu.assert_same("reset cells", st["cells"], fx["reset_cells"][ep], ctx)
        if wrapped:
            gu.assert_same("reset theta", st["theta"], fx["reset_theta"][ep], ctx)
        for t in range(fx["actions"].shape[1]):
            ctx = "ep %d step %d" % (ep, t)
            o = backend.step(fx["actions"][ep, t])
            st = backend.state()
            gu.assert_same("pos", st["pos"], fx["pos"][ep, t], ctx)
            gu.assert_same("ori", st["ori"], fx["ori"][ep, t], ctx)
            gu.assert_same("cells", st["cells"], fx["cells"][ep, t], ctx)
            gu.assert_same("obs", o["obs"], fx["obs"][ep, t][:, :F], ctx)
            gu.assert_same("rew", o["rew"], fx["rew"][ep, t], ctx)
            gu.assert_same("base_rew", o["base_rew"], fx["base_rew"][ep, t], ctx)
            if wrapped:
                gu.assert_same("transfers", o["transfers"], fx["transfers"][ep, t], ctx)
            gu.assert_same("info0", o["info"][:, 0], fx["info0"][ep, t], ctx)
            gu.assert_same("info1", o["info"][:, 1], fx["info1"][ep, t], ctx)
            gu.assert_same("done", int(o["done"]), int(fx["done"][ep, t]), ctx)
        raw = np.asarray(backend.metrics_raw(), dtype=np.float64)
        want = dict(zip([str(k) for k in fx["metric_keys"]], fx["metrics"][ep]))
        for k, v in _metrics(kind, n, raw).items():
            gu.assert_same("metric " + k, np.float64(v), np.float64(want[k]), "ep %d" % ep)
        if "equality" in want:
            from contracts_b200.environments.gridworld import equality, sustainability
            gu.assert_same("equality", np.float64(equality(list(raw[8:8 + n]))), np.float64(want["equality"]), "ep %d" % ep)
            gu.assert_same("sustainability", np.float64(sustainability(list(raw[8:8 + n]), list(raw[16:16 + n]))),
                           np.float64(want["sustainability"]), "ep %d" % ep)
            if wrapped:
                gu.assert_same("transfer_equality", np.float64(equality(list(raw[24:24 + n]))),
                               np.float64(want["transfer_equality"]), "ep %d" % ep)
                gu.assert_same("transfer_sustainability", np.float64(sustainability(list(raw[24:24 + n]), list(raw[32:32 + n]))),
                               np.float64(want["transfer_sustainability"]), "ep %d" % ep)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_features_match_reference(oracle_lib, name):
    replay(_Oracle(oracle_lib, gu.load(name)), gu.load(name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_features_match_reference(name):
    replay(_Cuda(gu.load(name)), gu.load(name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in NAMES if "nocontract" not in n])
def test_dropin_feature_env_dict_api(name):
    """env_creator('Cleanup' / 'Harvest') + ContractWrapperSubgame (non-convolutional): obs = concat(features, theta, [0])."""
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator, get_base_env_tag
    fx = gu.load(name)
    kind, n = str(fx["kind"]), int(fx["n"])
    base = env_creator(get_base_env_tag({"environment": kind}), dict(num_agents=n, horizon=int(fx["horizon"]),
                                                                      seed=int(fx["seed"]), env_id=int(fx["env_id"])))
    env = env_creator("ContractWrapperSubgame", dict(num_agents=n, base_env=base, contract=getattr(contract_list, gu.contract_name(fx))(n),
                                                     convolutional=False))
    keys = ["a%d" % i for i in range(n)]
    for ep in range(fx["actions"].shape[0]):
        obs = env.reset()
        gu.assert_same("reset obs", np.stack([obs[k] for k in keys]), fx["reset_obs"][ep], "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            ctx = "ep %d step %d" % (ep, t)
            obs, rew, done, info = env.step({k: int(a) for k, a in zip(keys, fx["actions"][ep, t])})
            gu.assert_same("obs", np.stack([obs[k] for k in keys]), fx["obs"][ep, t], ctx)
            gu.assert_same("rew", np.array([rew[k] for k in keys]), fx["rew"][ep, t], ctx)
            d = bool(fx["done"][ep, t])
            assert done == {"__all__": d, "a0": d, "a1": d}, ctx
        m = base.metrics
        want = dict(zip([str(k) for k in fx["metric_keys"]], fx["metrics"][ep]))
        assert set(m.keys()) == set(want.keys()), (sorted(m), sorted(want))
        for k, v in want.items():
            gu.assert_same("metric " + k, np.float64(m[k]), np.float64(v), "ep %d" % ep)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,E,nact", [("cleanup", 8, 2050, 9), ("harvest", 8, 1030, 8), ("cleanup", 3, 257, 8), ("harvest", 1, 64, 8)])
def test_cuda_features_rollout_matches_oracle(oracle_lib, kind, n, E, nact):
    """Random rollouts at batch size, every output incl. masked re-resets and metrics, bit-exact vs the C oracle."""
    import torch
    from contracts_b200.features import BatchedFeatureEnv
    contract = None if n < 2 else ("CleanupContract" if kind == "cleanup" else "HarvestFeaturemodLocalContract")
    env = BatchedFeatureEnv(kind, E, n, horizon=70, contract=contract, seed=9, first_env_id=4000000000)
    orc = oracle_lib.FeatOracle(kind, E, n, _maps(kind), horizon=70, contract=contract, seed=9, first_env_id=4000000000)
    gu.assert_same("reset obs", env.reset().cpu().numpy(), orc.reset(), "reset")
    rng = np.random.RandomState(n)
    p = None
    if kind == "cleanup":
        p = np.array([.11, .11, .11, .11, .05, .05, .05, .36, .05][:nact]); p = p / p.sum()
    for t in range(160):
        a = rng.choice(nact, size=(E, n), p=p)
        o = orc.step(a)
        obs, rew, done, info = env.step(torch.from_numpy(a.astype(np.uint8)).cuda())
        ctx = "step %d" % t
        gu.assert_same("obs", obs.cpu().numpy(), o["obs"], ctx)
        gu.assert_same("rew", rew.cpu().numpy(), o["rew"], ctx)
        gu.assert_same("base_rew", env.base_rew.cpu().numpy(), o["base_rew"], ctx)
        gu.assert_same("transfers", env.transfers.cpu().numpy(), o["transfers"], ctx)
        gu.assert_same("info", info.cpu().numpy()[..., :2], o["info"][..., :2], ctx)
        gu.assert_same("done", done.cpu().numpy(), o["done"], ctx)
        if o["done"].any():
            gu.assert_same("metrics", env.metrics_raw().cpu().numpy(), orc.metrics_raw(), ctx)
            mask = o["done"].copy(); mask[::4] = 0
            r1 = env.reset(torch.from_numpy(mask).cuda()).cpu().numpy()
            r2 = orc.reset(mask)
            gu.assert_same("masked reset obs", r1[mask.astype(bool)], r2[mask.astype(bool)], ctx)
    so, sc = orc.get_state(), env.get_state()
    for k in ("pos", "ori", "cells", "theta", "t"):
        gu.assert_same(k, sc[k].cpu().numpy(), so[k], "end")


@pytest.mark.gpu
@pytest.mark.parametrize("kind,contract,n", [("cleanup", "CleanupContract", 8), ("harvest", "HarvestFeaturemodLocalContract", 4)])
def test_feature_env_next_step_auto_reset(kind, contract, n):
    """ssd_feat_io.auto_reset: an env that reached its horizon is reset by the NEXT step call (reset observation, zero
    rewards, done cleared, actions ignored) — compared env by env with single-env handles driven by explicit reset()."""
    import torch
    from contracts_b200.features import BatchedFeatureEnv
    E, H, nact = 12, 9, 8 if kind == "cleanup" else 7
    batch = BatchedFeatureEnv(kind, E, n, horizon=H, contract=contract, seed=17, first_env_id=50)
    singles = [BatchedFeatureEnv(kind, 1, n, horizon=H, contract=contract, seed=17, first_env_id=50 + e) for e in range(E)]
    obs = batch.reset().cpu().numpy().copy()
    for e, s in enumerate(singles):
        assert np.array_equal(s.reset().cpu().numpy()[0], obs[e])
    # stagger: env e takes e % 4 extra steps before the comparison starts, so that the resets do not coincide
    rng = np.random.RandomState(5)
    done_prev = np.zeros(E, dtype=bool)
    for t in range(4 * H):
        acts = rng.randint(0, nact, size=(E, n)).astype(np.uint8)
        o, r, d, i = batch.step(torch.as_tensor(acts).cuda(), auto_reset=True)
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        for e, s in enumerate(singles):
            if done_prev[e]:
                so = s.reset().cpu().numpy()[0]
                assert np.array_equal(o[e], so) and not r[e].any() and d[e] == 0, (t, e)
            else:
                so, sr, sd, _ = s.step(torch.as_tensor(acts[e:e + 1]).cuda())
                assert np.array_equal(o[e].view(np.uint64), so.cpu().numpy()[0].view(np.uint64)), (t, e)
                assert np.array_equal(r[e].view(np.uint64), sr.cpu().numpy()[0].view(np.uint64)), (t, e)
                assert d[e] == int(sd[0].item())
        done_prev = d.astype(bool)
    assert t > 3 * H
