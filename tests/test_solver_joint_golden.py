"""CPU: the oracle's restatement of NegotiationSolver (candidates + decision rules) and of the JointEnv output
layouts against the reference-generated fixtures (oracle/make_golden.py: SOLVER_SCENARIOS, JOINT_SCENARIOS)."""
import numpy as np
import pytest

import golden_util as gu


def _grid(oracle_lib, fx, contract=True, **kw):
    from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP
    kind = str(fx["kind"])
    amap = CLEANUP_MAP if kind == "cleanup" else HARVEST_MAP
    cname = ("CleanupContract" if kind == "cleanup" else "HarvestFeaturemodLocalContract") if contract else None
    return oracle_lib.GridOracle(kind, 1, int(fx["n"]), amap, contract=cname, seed=int(fx["seed"]),
                                 first_env_id=int(fx["env_id"]), **kw)


@pytest.mark.parametrize("name", gu.fixture_names("solver_"))
def test_oracle_solver_matches_reference(oracle_lib, name):
    fx = gu.load(name)
    n, S, rule = int(fx["n"]), int(fx["num_samples"]), str(fx["rule"])
    orc = _grid(oracle_lib, fx)
    high = float(np.float32(0.2)) if str(fx["kind"]) == "cleanup" else float(np.float32(10.0))
    other = 0
    for ep in range(fx["params"].shape[0]):
        gu.assert_same("reset obs", orc.reset()[0], fx["reset_obs"][ep], "ep %d" % ep)
        params = oracle_lib.solver_candidates(int(fx["seed"]), int(fx["env_id"]), ep, 0.0, high, S)
        gu.assert_same("candidates", params, fx["params"][ep], "ep %d" % ep)
        theta, idx = oracle_lib.solver_choose(params, fx["vals"][ep], rule)
        gu.assert_same("theta", theta, fx["theta"][ep], "ep %d" % ep)
        other += oracle_lib.solver_choose(params, fx["vals"][ep], "max" if rule == "majority" else "majority")[1] != idx
        orc.set_theta(theta)
        for t in range(fx["actions"].shape[1]):
            o = orc.step(fx["actions"][ep, t][None], want_features=False)
            gu.assert_same("obs", o["obs"][0], fx["obs"][ep, t], "ep %d step %d" % (ep, t))
            gu.assert_same("rew", o["rew"][0], fx["rew"][ep, t], "ep %d step %d" % (ep, t))
            gu.assert_same("contract obs", np.array([theta, 0.0]), fx["contract_obs"][ep, t][0], "ep %d step %d" % (ep, t))
    if name in ("solver_cleanup_n4_majority", "solver_harvest_n8_majority"):
        assert other > 0, "fixture does not separate the two decision rules"


@pytest.mark.parametrize("name", gu.fixture_names("joint_"))
def test_oracle_joint_layouts_match_reference(oracle_lib, name):
    fx = gu.load(name)
    n, mode = int(fx["n"]), str(fx["mode"])
    orc = _grid(oracle_lib, fx, contract=False, horizon=int(fx["horizon"]))
    for ep in range(fx["actions"].shape[0]):
        obs = orc.reset()
        got = orc.global_view()[0] if mode == "global" else oracle_lib.concatenated_obs(obs)[0]
        gu.assert_same("reset obs", got, fx["reset_obs"][ep], "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            ctx = "ep %d step %d" % (ep, t)
            o = orc.step(fx["actions"][ep, t][None], want_features=True)
            got = orc.global_view()[0] if mode == "global" else oracle_lib.concatenated_obs(o["obs"])[0]
            gu.assert_same("obs", got, fx["obs"][ep, t], ctx)
            rew = 0                               # sum([rew for rew in env_rews.values()]) (two_stage_train.py:592)
            for r in o["rew"][0]:
                rew = rew + r
            gu.assert_same("rew", np.float64(rew), fx["rew"][ep, t], ctx)
            gu.assert_same("done", int(o["done"][0]), int(fx["done"][ep, t]), ctx)
            gu.assert_same("eaten_apples", int(o["info"][0, :, 0].sum()), int(fx["eaten_apples"][ep, t]), ctx)
            gu.assert_same("info1", int(o["info"][0, :, 1].sum()), int(fx["info1"][ep, t]), ctx)
            feat = 0
            for a in range(n):                    # sum([env_infos[agent]['feature_obs'] ...]) (:594-595)
                feat = feat + o["feature_obs"][0, a]
            gu.assert_same("feature_obs", feat, fx["feature_obs"][ep, t], ctx)


@pytest.mark.parametrize("name", gu.fixture_names("render_"))
def test_oracle_render_matches_reference(oracle_lib, name):
    """full_map_to_colors incl. the beams of the last step (map_env.py:354-375,389-392)."""
    fx = gu.load(name)
    orc = oracle_lib.GridOracle(str(fx["kind"]), 1, int(fx["n"]), [str(r) for r in fx["ascii_map"]], horizon=int(fx["horizon"]),
                                seed=int(fx["seed"]), first_env_id=int(fx["env_id"]))
    beams = 0
    for ep in range(fx["actions"].shape[0]):
        orc.reset()
        gu.assert_same("reset frame", orc.render()[0], fx["reset_obs"][ep], "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            orc.step(fx["actions"][ep, t][None], want_features=False)
            gu.assert_same("frame", orc.render()[0], fx["obs"][ep, t], "ep %d step %d" % (ep, t))
            beams += int(((fx["obs"][ep, t] == (255, 255, 0)).all(-1) | (fx["obs"][ep, t] == (100, 255, 255)).all(-1)).sum())
    assert beams > 20, "fixture shows no beams"
