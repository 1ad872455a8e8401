"""GPU: the reference's golden episodes replayed through the drop-in dict API
(`env_creator` tags, MultiAgentEnv reset()/step(), infos, dones, metrics) — what a user of
utils/env_creator_functions.py sees after switching packages."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _make(fx, image_obs=True):
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator, get_base_env_tag
    kind, n = str(fx["kind"]), int(fx["n"])
    tag = get_base_env_tag({"environment": kind + "_new"})
    base = env_creator(tag, dict(num_agents=n, env_params={}, image_obs=image_obs, disable_firing=False,
                                 ascii_map=[str(r) for r in fx["ascii_map"]], horizon=int(fx["horizon"]),
                                 seed=int(fx["seed"]), env_id=int(fx["env_id"]),
                                 return_agent_actions=bool(gu.shaping_kwargs(fx)),   # no effect, as in the reference
                                 **gu.shaping_kwargs(fx)))
    if not bool(fx["contract"]):
        return base, base
    cname = gu.contract_name(fx)
    env = env_creator("ContractWrapperSubgame", dict(num_agents=n, base_env=base, contract=getattr(contract_list, cname)(n),
                                                     convolutional=image_obs))
    return base, env


@pytest.mark.parametrize("name", gu.fixture_names())
def test_dict_api_replays_reference(name):
    fx = gu.load(name)
    kind, n = str(fx["kind"]), int(fx["n"])
    base, env = _make(fx)
    wrapped = bool(fx["contract"])
    keys = ["a%d" % i for i in range(n)]
    episodes, steps = fx["actions"].shape[:2]
    assert base.action_space.n == (9 if kind == "cleanup" else 8)
    assert set(env.observation_space.keys()) == ({"image", "contract"} if wrapped else {"image"})
    for ep in range(episodes):
        obs = env.reset()
        ctx = "reset ep %d" % ep
        assert sorted(obs.keys()) == keys
        for i, k in enumerate(keys):
            gu.assert_same("reset image", obs[k]["image"], fx["reset_obs"][ep][i].astype(np.float64) / 255, ctx)
            if wrapped:
                gu.assert_same("reset contract", obs[k]["contract"], np.array([fx["reset_theta"][ep], 0.0]), ctx)
        for t in range(steps):
            ctx = "ep %d step %d" % (ep, t)
            acts = {k: int(fx["actions"][ep, t][i]) for i, k in enumerate(keys)}
            obs, rew, done, info = env.step(acts)
            d = bool(fx["done"][ep, t])
            assert done == {"__all__": d, "a0": d, "a1": d}, ctx
            for i, k in enumerate(keys):
                gu.assert_same("image", obs[k]["image"], fx["obs"][ep, t][i].astype(np.float64) / 255, ctx)
                gu.assert_same("reward", rew[k], fx["rew"][ep, t][i], ctx)
                assert info[k]["eaten_apples"] == fx["eaten_apples"][ep, t][i], ctx
                assert info[k]["cleaned_squares" if kind == "cleanup" else "eaten_close_apples"] == fx["info1"][ep, t][i], ctx
                gu.assert_same("feature_obs", info[k]["feature_obs"], fx["feature_obs"][ep, t][i], ctx)
                if wrapped:
                    gu.assert_same("contract", obs[k]["contract"], np.array([fx["reset_theta"][ep], 0.0]), ctx)
                    gu.assert_same("contract_param", info[k]["contract_param"], np.array([fx["reset_theta"][ep]]), ctx)
                    assert isinstance(rew[k], np.float64)
                else:
                    assert isinstance(rew[k], int)
        m = env.metrics if not wrapped else base.metrics
        want = dict(zip([str(k) for k in fx["metric_keys"]], fx["metrics"][ep]))
        assert set(m.keys()) == set(want.keys()), "ep %d metric keys %s vs %s" % (ep, sorted(m), sorted(want))
        for k, v in want.items():
            gu.assert_same("metric " + k, np.float64(m[k]), np.float64(v), "ep %d" % ep)


def test_env_creator_rejects_unknown_and_out_of_scope():
    from contracts_b200.utils.env_creator_functions import TAGS, env_creator
    assert set(TAGS) == {"SelfDrive", "Harvest", "HarvestNew", "Cleanup", "CleanupNew", "ContractWrapperNegotiate",
                         "ContractWrapperSubgame", "ContractWrapperCombined", "NegotiationSolver", "JointEnv"}
    with pytest.raises(ValueError):
        env_creator("Nope", {})
    with pytest.raises(NotImplementedError):
        env_creator("ContractWrapperCombined", {})


def test_render_and_global_obs():
    from contracts_b200.utils.env_creator_functions import env_creator
    env = env_creator("CleanupNew", dict(num_agents=3, env_params={}))
    env.reset()
    img = env.render()
    assert img.shape == (25, 18, 3)
    st = env.batch.get_state()
    r, c = st["pos"][0, 2].tolist()
    assert tuple(img[r, c]) == (204, 0, 204)               # agent '3' colour (map_env.py:33)
    assert tuple(img[0, 0]) == (180, 180, 180)
    assert env.get_global_obs()["image"].max() <= 1.0


@pytest.mark.parametrize("name", gu.fixture_names("flatobs_"))
def test_flat_feature_observations(name):
    """image_obs=False + non-convolutional wrapper: obs = concat(feature vector, theta, [0]) (two_stage_train.py:113-121)."""
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    fx = gu.load(name)
    kind, n = str(fx["kind"]), int(fx["n"])
    base = env_creator("CleanupNew" if kind == "cleanup" else "HarvestNew",
                       dict(num_agents=n, env_params={}, image_obs=False, horizon=int(fx["horizon"]),
                            seed=int(fx["seed"]), env_id=int(fx["env_id"])))
    cname = "CleanupContract" if kind == "cleanup" else "HarvestFeaturemodLocalContract"
    env = env_creator("ContractWrapperSubgame", dict(num_agents=n, base_env=base, contract=getattr(contract_list, cname)(n),
                                                     convolutional=False))
    gu.assert_same("space low", env.observation_space.low, fx["space_low"], name)
    gu.assert_same("space high", env.observation_space.high, fx["space_high"], name)
    keys = ["a%d" % i for i in range(n)]
    for ep in range(fx["actions"].shape[0]):
        obs = env.reset()
        gu.assert_same("reset obs", np.stack([obs[k] for k in keys]), fx["reset_obs"][ep], "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            obs, rew, done, info = env.step({k: int(a) for k, a in zip(keys, fx["actions"][ep, t])})
            ctx = "ep %d step %d" % (ep, t)
            gu.assert_same("obs", np.stack([obs[k] for k in keys]), fx["obs"][ep, t], ctx)
            gu.assert_same("rew", np.array([rew[k] for k in keys]), fx["rew"][ep, t], ctx)


def test_seed_and_agent_pos():
    """env.seed(s) re-keys the env's random stream (MapEnv.seed, map_env.py:344-345); agent_pos lists [row, col] per agent."""
    from contracts_b200.utils.env_creator_functions import env_creator
    a = env_creator("HarvestNew", dict(num_agents=4, env_params={}, seed=5))
    b = env_creator("HarvestNew", dict(num_agents=4, env_params={}, seed=6))
    oa, ob = a.reset(), b.reset()
    assert any(not np.array_equal(oa[k]["image"], ob[k]["image"]) for k in oa) or a.agent_pos != b.agent_pos
    assert b.seed(5) == [5]
    ob = b.reset()
    for k in oa:
        gu.assert_same("image after reseeding", ob[k]["image"], oa[k]["image"], k)
    assert a.agent_pos == b.agent_pos and len(a.agent_pos) == 4 and len(a.agent_pos[0]) == 2
