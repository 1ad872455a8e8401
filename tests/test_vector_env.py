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
