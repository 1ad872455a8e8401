"""SeparateContractNegotiateStage (two_stage_train.py:190-358): the reference's golden episodes
(tests/golden/negotiate_*.npz: reset -> proposal -> agreement + scripted frozen-policy rollout) replayed
through the C oracle on CPU and through the drop-in wrapper (CUDA) on GPU."""
import numpy as np
import pytest

import golden_util as gu

NAMES = gu.fixture_names("negotiate_")


def test_negotiate_fixtures_present():
    assert len(NAMES) >= 3


def _contract_name(kind):
    return "CleanupContract" if kind == "cleanup" else "HarvestFeaturemodLocalContract"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_negotiation_matches_reference(oracle_lib, name):
    from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP
    fx = gu.load(name)
    kind, n = str(fx["kind"]), int(fx["n"])
    amap = CLEANUP_MAP if kind == "cleanup" else HARVEST_MAP
    o = oracle_lib.GridOracle(kind, 1, n, amap, horizon=int(fx["base_horizon"]), contract=_contract_name(kind),
                              seed=int(fx["seed"]), first_env_id=int(fx["env_id"]))
    steps = min(int(fx["horizon"]), int(fx["base_horizon"]))
    for ep in range(fx["acts"].shape[0]):
        ctx = "ep %d" % ep
        gu.assert_same("reset obs", o.reset()[0], fx["reset_obs"][ep], ctx)
        acts = fx["acts"][ep]
        dec = o.negotiate(acts[0, 0], acts[:, 1])
        gu.assert_same("accepted", dec[0], fx["accepted"][ep], ctx)
        theta = o.get_state()["theta"][0]
        gu.assert_same("theta", theta, fx["contract_obs3"][ep][0, 0], ctx)
        total = np.zeros(n)
        for t in range(steps):
            r = o.step(fx["table"][t][None], want_features=False)
            total = total + r["rew"][0]
        gu.assert_same("summed rewards", total, fx["rew3"][ep], ctx)
        gu.assert_same("final obs", r["obs"][0], fx["obs3"][ep], ctx)
        assert int(o.get_state()["t"][0]) == int(fx["t_end"][ep])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_dropin_negotiate_stage_matches_reference(name):
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    fx = gu.load(name)
    kind, n = str(fx["kind"]), int(fx["n"])
    base = env_creator("CleanupNew" if kind == "cleanup" else "HarvestNew",
                       dict(num_agents=n, env_params={}, image_obs=True, horizon=int(fx["base_horizon"]),
                            seed=int(fx["seed"]), env_id=int(fx["env_id"])))
    contract = getattr(contract_list, _contract_name(kind))(n)
    calls = {"i": 0}
    table = fx["table"]

    def policy(obs, key):                      # the scripted stand-in for the frozen PPO policy
        assert set(obs.keys()) >= {"image", "contract"}
        t, i = divmod(calls["i"], n)
        calls["i"] += 1
        return int(table[t % table.shape[0], i])

    env = env_creator("ContractWrapperNegotiate", dict(base_env=base, contract=contract, num_agents=n,
                                                       horizon=int(fx["horizon"]), trainer_config=None, trainer_env=None,
                                                       trainer_path=None, convolutional=True, shared=True, policy=policy))
    keys = ["a%d" % i for i in range(n)]

    def check(obs, want_img, want_contract, ctx):
        for i, k in enumerate(keys):
            gu.assert_same("image", obs[k]["image"], want_img[i].astype(np.float64) / 255, ctx)
            gu.assert_same("contract", obs[k]["contract"], want_contract[i], ctx)

    for ep in range(fx["acts"].shape[0]):
        calls["i"] = 0
        ctx = "ep %d" % ep
        check(env.reset(), fx["reset_obs"][ep], fx["reset_contract_obs"][ep], ctx + " reset")
        acts = {k: fx["acts"][ep][i] for i, k in enumerate(keys)}
        obs, rew, done, info = env.step(acts)
        assert done == {"__all__": False}
        check(obs, fx["obs2"][ep], fx["contract_obs2"][ep], ctx + " proposal")
        gu.assert_same("rew2", [rew[k] for k in keys], fx["rew2"][ep], ctx)
        obs, rew, done, info = env.step(acts)
        assert done == {"__all__": True}
        assert env.metrics["accepted"] == int(fx["accepted"][ep])
        check(obs, fx["obs3"][ep], fx["contract_obs3"][ep], ctx + " agreement")
        gu.assert_same("rew3", [rew[k] for k in keys], fx["rew3"][ep], ctx)
        m = base.metrics
        for k, v in zip([str(x) for x in fx["metric_keys"]], fx["base_metrics"][ep]):
            gu.assert_same("metric " + k, np.float64(m[k]), np.float64(v), ctx)
