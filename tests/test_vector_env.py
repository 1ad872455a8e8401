"""GPU: the BaseEnv-shaped vector adapter returns, env by env, what the single-env drop-in classes return."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_vector_env_matches_single_env_facade():
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    from contracts_b200.vector_env import SSDVectorEnv
    n, E, H = 4, 6, 12
    vec = SSDVectorEnv("cleanup_new", E, n, contract="CleanupContract", horizon=H, seed=7, first_env_id=100)
    singles = []
    for e in range(E):
        base = env_creator("CleanupNew", dict(num_agents=n, env_params={}, image_obs=True, horizon=H, seed=7, env_id=100 + e))
        singles.append(env_creator("ContractWrapperSubgame", dict(num_agents=n, base_env=base, contract=contract_list.CleanupContract(n),
                                                                  convolutional=True)))
    keys = ["a%d" % i for i in range(n)]
    want = [s.reset() for s in singles]
    obs, rews, dones, infos, _ = vec.poll()
    assert rews == {} and dones == {}
    rng = np.random.RandomState(3)

    def same_obs(got, exp, ctx):
        for k in keys:
            assert np.array_equal(got[k]["image"], exp[k]["image"]), ctx
            assert np.array_equal(got[k]["contract"], exp[k]["contract"]), ctx

    for e in range(E):
        same_obs(obs[e], want[e], ("reset", e))
    for t in range(2 * H + 3):
        acts = {e: {k: int(rng.randint(9)) for k in keys} for e in range(E)}
        vec.send_actions(acts)
        obs, rews, dones, infos, _ = vec.poll()
        for e in range(E):
            o, r, d, i = singles[e].step(acts[e])
            same_obs(obs[e], o, (t, e))
            assert all(np.float64(rews[e][k]).tobytes() == np.float64(r[k]).tobytes() for k in keys), (t, e)
            assert dones[e]["__all__"] == d["__all__"]
            assert all(infos[e][k]["cleaned_squares"] == i[k]["cleaned_squares"] for k in keys)
            if d["__all__"]:
                vec.try_reset(e)
                want[e] = singles[e].reset()
        if any(dones[e]["__all__"] for e in range(E)):
            obs, rews, dones, infos, _ = vec.poll()              # the reset observations; no rewards / dones for fresh envs
            for e in range(E):
                same_obs(obs[e], want[e], ("re-reset", t, e))
                assert e not in rews
    vec.stop()


@pytest.mark.parametrize("kind,tag,cname", [("cleanup", "Cleanup", "CleanupContract"), ("harvest", "Harvest", "HarvestFeaturemodLocalContract")])
def test_feature_vector_env_matches_single_env_facade(kind, tag, cname):
    """SSDFeatureVectorEnv.poll() / send_actions() / try_reset() env by env against env_creator('Cleanup' / 'Harvest') +
    ContractWrapperSubgame (non-convolutional: observation = features ++ [theta, 0])."""
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    from contracts_b200.vector_env import SSDFeatureVectorEnv
    n, E, H = 4, 5, 9
    vec = SSDFeatureVectorEnv(kind, E, n, contract=cname, horizon=H, seed=7, first_env_id=100)
    singles = []
    for e in range(E):
        base = env_creator(tag, dict(num_agents=n, horizon=H, seed=7, env_id=100 + e))
        singles.append(env_creator("ContractWrapperSubgame", dict(num_agents=n, base_env=base, contract=getattr(contract_list, cname)(n),
                                                                  convolutional=False)))
    keys = ["a%d" % i for i in range(n)]
    want = [s.reset() for s in singles]
    obs, rews, dones, infos, _ = vec.poll()
    assert rews == {} and dones == {}
    for e in range(E):
        for k in keys:
            assert np.array_equal(np.asarray(obs[e][k], dtype=np.float64), np.asarray(want[e][k], dtype=np.float64)), ("reset", e, k)
    rng = np.random.RandomState(5)
    nact = 9 if kind == "cleanup" else 8
    for t in range(2 * H + 2):
        acts = {e: {k: int(rng.randint(nact)) for k in keys} for e in range(E)}
        vec.send_actions(acts)
        obs, rews, dones, infos, _ = vec.poll()
        for e in range(E):
            o, r, d, _i = singles[e].step(acts[e])
            for k in keys:
                assert np.array_equal(np.asarray(obs[e][k], dtype=np.float64), np.asarray(o[k], dtype=np.float64)), (t, e, k)
                assert np.float64(rews[e][k]).tobytes() == np.float64(r[k]).tobytes(), (t, e, k)
            assert dones[e]["__all__"] == d["__all__"]
            if d["__all__"]:
                vec.try_reset(e)
                want[e] = singles[e].reset()
        if any(dones[e]["__all__"] for e in range(E)):
            obs, rews, dones, infos, _ = vec.poll()
            for e in obs:
                for k in keys:
                    assert np.array_equal(np.asarray(obs[e][k], dtype=np.float64), np.asarray(want[e][k], dtype=np.float64)), ("re-reset", e, k)
            assert rews == {}


def test_car_vector_env_arrays_match_batch():
    """SSDCarVectorEnv: array polls equal a plain BatchedCarEnv driven the same way (masked resets on done)."""
    import torch
    from contracts_b200.selfdrive import BatchedCarEnv
    from contracts_b200.vector_env import SSDCarVectorEnv
    n, E = 4, 12
    vec = SSDCarVectorEnv(E, n, contract="SelfdriveContractDistprop", seed=3, first_env_id=50)
    ref = BatchedCarEnv(E, n, contract="SelfdriveContractDistprop", seed=3, first_env_id=50)
    s = vec.poll_arrays()
    assert np.array_equal(s["obs"], ref.reset().cpu().numpy()) and s["fresh"].all()
    rng = np.random.RandomState(1)
    resets = 0
    for t in range(400):
        a = (rng.uniform(-0.02, 0.1, size=(E, n))).astype(np.float32)
        vec.send_action_array(a)
        s = vec.poll_arrays()
        obs, rew, done, _ = ref.step(torch.from_numpy(a).cuda())
        assert np.array_equal(s["obs"].view(np.uint64), obs.cpu().numpy().view(np.uint64)), t
        assert np.array_equal(s["rew"].view(np.uint64), rew.cpu().numpy().view(np.uint64)), t
        assert np.array_equal(s["done"], done.cpu().numpy()), t
        fin = s["done"][:, n].astype(bool)
        if fin.any():
            for e in np.nonzero(fin)[0]:
                vec.try_reset(int(e))
            ref.reset(torch.from_numpy(fin.astype(np.uint8)).cuda())
            resets += int(fin.sum())
            s = vec.poll_arrays()
            assert np.array_equal(s["obs"][fin], ref.obs.cpu().numpy()[fin]) and np.array_equal(s["fresh"], fin)
    assert resets >= E
