"""GPU: NegotiationSolver (candidates + decision rules) and the JointEnv output layouts on the device, against the
reference-generated fixtures (batched C-ABI path and drop-in dict API) and against the oracle at larger sizes."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _batch(fx, E=3, index=1, contract=True, **kw):
    from contracts_b200.batched import BatchedGridEnv
    kind = str(fx["kind"])
    cname = ("CleanupContract" if kind == "cleanup" else "HarvestFeaturemodLocalContract") if contract else None
    return BatchedGridEnv(kind + "_new", E, int(fx["n"]), contract=cname, seed=int(fx["seed"]),
                          first_env_id=(int(fx["env_id"]) - index) & 0xFFFFFFFF, **kw)


@pytest.mark.parametrize("name", gu.fixture_names("solver_"))
def test_solver_batched_matches_reference(name):
    import torch
    fx = gu.load(name)
    n, S, rule, i = int(fx["n"]), int(fx["num_samples"]), str(fx["rule"]), 1
    env = _batch(fx)
    rng = np.random.RandomState(5)
    for ep in range(fx["params"].shape[0]):
        gu.assert_same("reset obs", env.reset()[i].cpu().numpy(), fx["reset_obs"][ep], "ep %d" % ep)
        params = env.solver_sample(S)
        gu.assert_same("candidates", params[i].cpu().numpy(), fx["params"][ep], "ep %d" % ep)
        vals = rng.uniform(-1, 1, size=(env.E, S + 1, n))
        vals[i] = fx["vals"][ep]
        theta, idx = env.solver_choose(params, torch.as_tensor(vals).cuda(), rule)
        gu.assert_same("theta", theta[i].item(), fx["theta"][ep], "ep %d" % ep)
        gu.assert_same("theta in state", env.get_state()["theta"].cpu().numpy(), theta.cpu().numpy(), "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            a = torch.as_tensor(np.broadcast_to(fx["actions"][ep, t].astype(np.uint8), (env.E, n)).copy()).cuda()
            obs, rew, done, info = env.step(a)
            gu.assert_same("obs", obs[i].cpu().numpy(), fx["obs"][ep, t], "ep %d step %d" % (ep, t))
            gu.assert_same("rew", rew[i].cpu().numpy(), fx["rew"][ep, t], "ep %d step %d" % (ep, t))


@pytest.mark.parametrize("name", gu.fixture_names("solver_"))
def test_solver_dict_api_replays_reference(name):
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    from oracle.scripted import scripted_value
    fx = gu.load(name)
    kind, n, S, rule = str(fx["kind"]), int(fx["n"]), int(fx["num_samples"]), str(fx["rule"])
    base = env_creator("CleanupNew" if kind == "cleanup" else "HarvestNew",
                       dict(num_agents=n, env_params={}, image_obs=True, seed=int(fx["seed"]), env_id=int(fx["env_id"])))
    contract = getattr(contract_list, gu.contract_name({"contract": True, "kind": kind}))(n)
    scale = float(contract.contract_space.high[0])
    env = env_creator("NegotiationSolver", dict(
        base_env=base, contract=contract, num_agents=n, horizon=1000, trainer_config={}, trainer_env=None, trainer_path=None,
        convolutional=True, shared=True, contract_samples=S, decision_rule=rule,
        value_fn=lambda obs, k: scripted_value(obs, int(k[1:]), n, scale)))
    keys = ["a%d" % i for i in range(n)]
    for ep in range(fx["params"].shape[0]):
        obs = env.reset()
        for i, k in enumerate(keys):
            gu.assert_same("reset image", obs[k]["image"], fx["reset_obs"][ep][i].astype(np.float64) / 255, "ep %d" % ep)
            gu.assert_same("reset contract", obs[k]["contract"], fx["reset_contract_obs"][ep][i], "ep %d" % ep)
        gu.assert_same("theta", np.float64(env.contract_param[0]), fx["theta"][ep], "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            ctx = "ep %d step %d" % (ep, t)
            obs, rew, done, info = env.step({k: int(fx["actions"][ep, t][i]) for i, k in enumerate(keys)})
            for i, k in enumerate(keys):
                gu.assert_same("image", obs[k]["image"], fx["obs"][ep, t][i].astype(np.float64) / 255, ctx)
                gu.assert_same("contract", obs[k]["contract"], fx["contract_obs"][ep, t][i], ctx)
                gu.assert_same("reward", rew[k], fx["rew"][ep, t][i], ctx)


def test_solver_rules_match_oracle_with_ties(oracle_lib):
    """Random values on a coarse lattice (many welfare ties and duplicate candidates), both rules, every env kind."""
    import torch
    from contracts_b200.batched import BatchedGridEnv
    from contracts_b200.features import BatchedFeatureEnv
    from contracts_b200.selfdrive import BatchedCarEnv
    rng = np.random.RandomState(11)
    E, S = 300, 7
    envs = [(BatchedGridEnv("cleanup_new", E, 5, contract="CleanupContract", seed=9, first_env_id=1000), float(np.float32(0.2)), 1000),
            (BatchedFeatureEnv("harvest", E, 4, contract="HarvestFeaturemodLocalContract", seed=9, first_env_id=7), float(np.float32(10.0)), 7),
            (BatchedCarEnv(E, 3, contract="SelfdriveContractDistprop", seed=9, first_env_id=50), float(np.float32(100.0)), 50)]
    from contracts_b200.batched import solver_choose, solver_sample
    for env, high, first in envs:
        n = env.n
        for ep in range(2):
            env.reset()
            params = solver_sample(env, S)
            pc = params.cpu().numpy()
            for e in (0, 1, E - 1):
                want = oracle_lib.solver_candidates(9, first + e, ep, 0.0, high, S)
                gu.assert_same("candidates", pc[e], want, "%s ep %d env %d" % (type(env).__name__, ep, e))
            for rule in ("max", "majority"):
                vals = rng.randint(0, 3, size=(E, S + 1, n)).astype(np.float64) * 0.5
                vals[::5, 3] = vals[::5, 1]                 # duplicate candidates: all_vals.index() picks the first
                theta, idx = solver_choose(env, params, torch.as_tensor(vals).cuda(), rule)
                th, ix = theta.cpu().numpy(), idx.cpu().numpy()
                for e in range(E):
                    wt, wi = oracle_lib.solver_choose(pc[e], vals[e], rule)
                    assert ix[e] == wi and gu.bits(th[e]) == gu.bits(wt), (type(env).__name__, rule, e, ix[e], wi)


@pytest.mark.parametrize("name", gu.fixture_names("joint_"))
def test_joint_layouts_batched_match_reference(name):
    import torch
    fx = gu.load(name)
    n, mode, i = int(fx["n"]), str(fx["mode"]), 2
    env = _batch(fx, E=5, index=i, contract=False, horizon=int(fx["horizon"]))
    view = (lambda: env.global_view()[i].cpu().numpy()) if mode == "global" else (lambda: env.concatenated_obs()[i].cpu().numpy())
    for ep in range(fx["actions"].shape[0]):
        env.reset()
        gu.assert_same("reset obs", view(), fx["reset_obs"][ep], "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            a = torch.as_tensor(np.broadcast_to(fx["actions"][ep, t].astype(np.uint8), (env.E, n)).copy()).cuda()
            obs, rew, done, info = env.step(a)
            gu.assert_same("obs", view(), fx["obs"][ep, t], "ep %d step %d" % (ep, t))
            gu.assert_same("done", int(done[i].item()), int(fx["done"][ep, t]), "ep %d step %d" % (ep, t))


@pytest.mark.parametrize("name", gu.fixture_names("joint_"))
def test_joint_dict_api_replays_reference(name):
    from contracts_b200.utils.env_creator_functions import env_creator
    fx = gu.load(name)
    kind, n, mode = str(fx["kind"]), int(fx["n"]), str(fx["mode"])
    base = env_creator("CleanupNew" if kind == "cleanup" else "HarvestNew",
                       dict(num_agents=n, env_params={}, image_obs=True, disable_firing=False, horizon=int(fx["horizon"]),
                            seed=int(fx["seed"]), env_id=int(fx["env_id"])))
    env = env_creator("JointEnv", dict(base_env=base, num_agents=n, global_obs=mode == "global",
                                       concatenated_obs=mode == "concatenated"))
    assert tuple(env.observation_space["image"].shape) == tuple(fx["reset_obs"].shape[1:])
    assert list(env.action_space.nvec) == [9 if kind == "cleanup" else 8] * n
    for ep in range(fx["actions"].shape[0]):
        obs = env.reset()
        assert list(obs.keys()) == ["a0"]
        gu.assert_same("reset image", obs["a0"]["image"], fx["reset_obs"][ep].astype(np.float64) / 255, "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            ctx = "ep %d step %d" % (ep, t)
            obs, rew, done, info = env.step({"a0": fx["actions"][ep, t]})
            gu.assert_same("image", obs["a0"]["image"], fx["obs"][ep, t].astype(np.float64) / 255, ctx)
            gu.assert_same("reward", np.float64(rew["a0"]), fx["rew"][ep, t], ctx)
            d = bool(fx["done"][ep, t])
            assert done == {"a0": d, "__all__": d}, ctx
            assert info["a0"]["eaten_apples"] == fx["eaten_apples"][ep, t], ctx
            assert info["a0"]["cleaned_squares" if kind == "cleanup" else "eaten_close_apples"] == fx["info1"][ep, t], ctx
            gu.assert_same("feature_obs", info["a0"]["feature_obs"], fx["feature_obs"][ep, t], ctx)


@pytest.mark.parametrize("kind,n,E", [("cleanup", 8, 257), ("harvest", 3, 130), ("cleanup", 1, 6)])
def test_joint_layouts_match_oracle(oracle_lib, kind, n, E):
    """Larger batches (E not a multiple of the kernel's 4-env groups), dense and padded observation tensors."""
    import torch
    from contracts_b200.batched import BatchedGridEnv
    from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP
    amap = CLEANUP_MAP if kind == "cleanup" else HARVEST_MAP
    for padded in (False, True):
        orc = oracle_lib.GridOracle(kind, E, n, amap, seed=3, first_env_id=10)
        env = BatchedGridEnv(kind + "_new", E, n, amap, seed=3, first_env_id=10, padded_obs=padded)
        rng = np.random.RandomState(2)
        obs_o = orc.reset()
        env.reset()
        for t in range(12):
            gu.assert_same("global view", env.global_view().cpu().numpy(), orc.global_view(), "t %d" % t)
            gu.assert_same("concatenated", env.concatenated_obs().cpu().numpy(), oracle_lib.concatenated_obs(obs_o), "t %d" % t)
            a = rng.randint(0, 8, size=(E, n))
            obs_o = orc.step(a, want_features=False)["obs"]
            env.step(torch.as_tensor(a.astype(np.uint8)).cuda())


@pytest.mark.parametrize("kind,n,E,padded", [("cleanup", 8, 37, False), ("harvest", 3, 20, True), ("cleanup", 5, 9, False)])
def test_policy_inputs_match_vision_net_preprocessing(kind, n, E, padded):
    """image = (curr_obs / 255 in float64 -> float32).permute(0, 3, 1, 2), contract = (theta, 0) x 5
    (environments/Networks/vision_net.py:159-167), in float32 / float16 / bfloat16."""
    import torch
    from contracts_b200.batched import BatchedGridEnv
    cname = "CleanupContract" if kind == "cleanup" else "HarvestFeaturemodLocalContract"
    env = BatchedGridEnv(kind + "_new", E, n, contract=cname, seed=4, first_env_id=3, padded_obs=padded)
    env.reset()
    rng = np.random.RandomState(1)
    for t in range(5):
        env.step(torch.as_tensor(rng.randint(0, 8, size=(E, n)).astype(np.uint8)).cuda())
    obs = env.obs.cpu().numpy()
    want = torch.from_numpy((obs / 255).reshape(E * n, 15, 15, 3)).float().permute(0, 3, 1, 2).contiguous()
    theta = env.get_state()["theta"].cpu()
    wc = torch.stack([theta, torch.zeros_like(theta)], dim=1).float().repeat_interleave(n, dim=0).repeat(1, 5)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        img, con = env.policy_inputs(dt)
        assert img.dtype == dt and tuple(img.shape) == (E * n, 3, 15, 15) and tuple(con.shape) == (E * n, 10)
        assert torch.equal(img.cpu(), want.to(dt)), dt
        assert torch.equal(con.cpu(), wc.to(dt)), dt
    assert torch.equal(env.policy_inputs(torch.float32, with_contract=False).cpu(), want)


@pytest.mark.parametrize("name", gu.fixture_names("render_"))
def test_render_with_beams_matches_reference(logic_variant, name):
    """full_map_to_colors incl. the beams of the last step (map_env.py:354-375,389-392): batched path and dict API."""
    import torch
    from contracts_b200.batched import BatchedGridEnv
    from contracts_b200.utils.env_creator_functions import env_creator
    fx = gu.load(name)
    kind, n, i = str(fx["kind"]), int(fx["n"]), 1
    amap = [str(r) for r in fx["ascii_map"]]
    env = BatchedGridEnv(kind + "_new", 6, n, amap, horizon=int(fx["horizon"]), seed=int(fx["seed"]),
                         first_env_id=(int(fx["env_id"]) - i) & 0xFFFFFFFF)
    env.record_beams()
    dropin = env_creator("CleanupNew" if kind == "cleanup" else "HarvestNew",
                         dict(num_agents=n, env_params={}, ascii_map=amap, horizon=int(fx["horizon"]), disable_firing=False,
                              seed=int(fx["seed"]), env_id=int(fx["env_id"])))
    keys = ["a%d" % k for k in range(n)]
    for ep in range(fx["actions"].shape[0]):
        env.reset()
        dropin.reset()
        gu.assert_same("reset frame", env.render()[i].cpu().numpy(), fx["reset_obs"][ep], "ep %d" % ep)
        gu.assert_same("reset frame (dict API)", dropin.render(mode="rgb_array"), fx["reset_obs"][ep], "ep %d" % ep)
        for t in range(fx["actions"].shape[1]):
            a = torch.as_tensor(np.broadcast_to(fx["actions"][ep, t].astype(np.uint8), (env.E, n)).copy()).cuda()
            env.step(a)
            dropin.step({k: int(fx["actions"][ep, t][j]) for j, k in enumerate(keys)})
            gu.assert_same("frame", env.render()[i].cpu().numpy(), fx["obs"][ep, t], "ep %d step %d" % (ep, t))
            gu.assert_same("frame (dict API)", dropin.full_map_to_colors(), fx["obs"][ep, t], "ep %d step %d" % (ep, t))


def test_render_matches_oracle_masked_reset(logic_variant, oracle_lib):
    """Beam overlays across a larger batch, cleared per env by masked resets."""
    import torch
    from contracts_b200.batched import BatchedGridEnv
    from contracts_b200.maps import CLEANUP_MAP
    E, n = 131, 8
    orc = oracle_lib.GridOracle("cleanup", E, n, CLEANUP_MAP, horizon=15, seed=5, first_env_id=2)
    env = BatchedGridEnv("cleanup_new", E, n, CLEANUP_MAP, horizon=15, seed=5, first_env_id=2)
    env.record_beams()
    rng = np.random.RandomState(3)
    orc.reset(); env.reset()
    for t in range(40):
        a = rng.choice(9, size=(E, n), p=[.1, .1, .1, .1, .05, .1, .1, .2, .15])
        o = orc.step(a, want_features=False)
        env.step(torch.as_tensor(a.astype(np.uint8)).cuda())
        gu.assert_same("frame", env.render().cpu().numpy(), orc.render(), "t %d" % t)
        if o["done"].any():
            mask = o["done"].copy(); mask[::2] = 0
            orc.reset(mask); env.reset(torch.as_tensor(mask).cuda())
            gu.assert_same("frame after masked reset", env.render().cpu().numpy(), orc.render(), "t %d" % t)


def test_joint_env_flat_variant_over_selfdrive():
    """JointEnv without image observations (duplicate_obs, Box action spaces; two_stage_train.py:509-570) over the
    selfdrive env: the joint step must equal the per-agent step of an identically seeded env, concatenated / summed."""
    from contracts_b200.utils.env_creator_functions import env_creator
    n = 4
    solo = env_creator("SelfDrive", dict(num_agents=n, seed=7, env_id=3))
    base = env_creator("SelfDrive", dict(num_agents=n, seed=7, env_id=3))
    joint = env_creator("JointEnv", dict(base_env=base, num_agents=n, duplicate_obs=True))
    keys = ["a%d" % i for i in range(n)]
    assert joint.observation_space.shape == (n * solo.observation_space.shape[0],) and joint.action_space.shape == (n,)
    o_solo, o_joint = solo.reset(), joint.reset()
    gu.assert_same("reset obs", o_joint["a0"], np.concatenate([o_solo[k] for k in keys]), "reset")
    rng = np.random.RandomState(4)
    last = dict(o_solo)
    active = list(keys)
    for t in range(60):
        a = rng.uniform(-0.1, 0.1, size=n).astype(np.float32)
        o_s, r_s, d_s, i_s = solo.step({k: np.array([a[i]]) for i, k in enumerate(keys) if k in active})
        o_j, r_j, d_j, i_j = joint.step({"a0": a})
        last.update(o_s)
        gu.assert_same("obs", o_j["a0"], np.concatenate([last[k] for k in keys]), "t %d" % t)
        gu.assert_same("rew", np.float64(r_j["a0"]), np.float64(sum([r for r in r_s.values()])), "t %d" % t)
        assert d_j == {"a0": d_s["__all__"], "__all__": d_s["__all__"]}
        active = [k for k in active if not d_s.get(k, False)]
        if d_s["__all__"]:
            break
    assert t > 5


@pytest.mark.parametrize("tag,contract_name,kwargs", [
    ("SelfDrive", "SelfdriveContractDistprop", dict(num_agents=4)),
    ("Cleanup", "CleanupContract", dict(num_agents=3, env_params={})),
    ("Harvest", "HarvestFeaturemodLocalContract", dict(num_agents=3, env_params={})),
])
def test_negotiation_solver_over_flat_envs(oracle_lib, tag, contract_name, kwargs):
    """NegotiationSolver (non-convolutional observations) over the selfdrive and feature envs: the contract it installs is
    the oracle's choice for the same candidates and values, and it shows up in the observation tail (two_stage_train.py:684-690)."""
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    n = kwargs["num_agents"]
    base = env_creator(tag, dict(seed=31, env_id=9, **kwargs))
    contract = getattr(contract_list, contract_name)(n)
    low, high = float(contract.contract_space.low[0]), float(contract.contract_space.high[0])
    seen = []

    def value_fn(obs, k):                      # depends on the candidate contract only: obs[-2] = theta
        theta = float(obs[-2]) / high
        i = int(k[1:])
        v = np.float64(1.0 + (0.3 + 0.2 * i) * theta - (0.5 + 0.4 * i) * theta * theta)
        seen.append((float(obs[-2]), i, v))
        return v

    env = env_creator("NegotiationSolver", dict(base_env=base, contract=contract, num_agents=n, horizon=100, trainer_config={},
                                                trainer_env=None, trainer_path=None, convolutional=False, shared=True,
                                                contract_samples=9, decision_rule="majority", value_fn=value_fn))
    for ep in range(2):
        seen.clear()
        obs = env.reset()
        params = oracle_lib.solver_candidates(31, 9, ep, low, high, 9)
        vals = np.array([v for _, _, v in seen]).reshape(10, n)
        gu.assert_same("candidates seen by the value function", np.array([t for t, i, _ in seen if i == 0]), params, "ep %d" % ep)
        want, _ = oracle_lib.solver_choose(params, vals, "majority")
        gu.assert_same("chosen contract", np.float64(env.contract_param[0]), want, "ep %d" % ep)
        for k in obs:
            gu.assert_same("observation tail", obs[k][-2:], np.array([want, 0.0]), "ep %d %s" % (ep, k))


def test_negotiate_every_env_kind_matches_oracle(oracle_lib):
    """ssd_negotiate on feature / selfdrive handles: the agreement draw depends only on (seed, env id, episode), so the
    grid oracle with the same keys gives the expected decisions; theta = proposal or 0 (two_stage_train.py:266-281)."""
    import torch
    from contracts_b200.batched import negotiate
    from contracts_b200.features import BatchedFeatureEnv
    from contracts_b200.maps import CLEANUP_MAP
    from contracts_b200.selfdrive import BatchedCarEnv
    E = 500
    rng = np.random.RandomState(3)
    for n, make in ((5, lambda: BatchedCarEnv(E, 5, contract="SelfdriveContractDistprop", seed=77, first_env_id=1000)),
                    (3, lambda: BatchedFeatureEnv("cleanup", E, 3, contract="CleanupContract", seed=77, first_env_id=1000)),
                    (8, lambda: BatchedFeatureEnv("harvest", E, 8, contract="HarvestFeaturemodLocalContract", seed=77, first_env_id=1000))):
        env = make()
        orc = oracle_lib.GridOracle("cleanup", E, n, CLEANUP_MAP, contract="CleanupContract", seed=77, first_env_id=1000)
        for ep in range(2):
            env.reset(); orc.reset()
            prop = rng.uniform(0, 0.2, size=E)
            acc = rng.uniform(0.3, 1.0, size=(E, n))
            dec = negotiate(env, torch.as_tensor(prop), torch.as_tensor(acc)).cpu().numpy()
            want = orc.negotiate(prop, acc)
            gu.assert_same("decisions", dec, want, "%s ep %d" % (type(env).__name__, ep))
            assert 0 < dec.sum() < E
            theta = env.get_state()["theta"].cpu().numpy()
            gu.assert_same("theta", theta, np.where(want == 1, prop, 0.0), "%s ep %d" % (type(env).__name__, ep))


def test_negotiate_stage_dict_api_over_selfdrive():
    """ContractWrapperNegotiate over the selfdrive env with a scripted frozen policy: propose -> agree -> rollout."""
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    n = 3
    base = env_creator("SelfDrive", dict(num_agents=n, seed=8, env_id=2))
    env = env_creator("ContractWrapperNegotiate", dict(
        base_env=base, contract=contract_list.SelfdriveContractDistprop(n), num_agents=n, horizon=400, convolutional=False,
        policy=lambda obs, k: np.array([0.1], dtype=np.float32)))
    keys = ["a%d" % i for i in range(n)]
    obs = env.reset()
    assert all(obs[k][-1] == 2 for k in keys)                           # negotiation state flag (two_stage_train.py:246-255)
    acts = {k: np.array([50.0, 1.0]) for k in keys}                     # proposal 50, everybody accepts
    obs, rew, done, info = env.step(acts)
    assert done == {"__all__": False} and all(obs[k][-1] == 3 for k in keys)
    obs, rew, done, info = env.step(acts)
    assert done["__all__"] and env.metrics["accepted"] == 1
    assert all(np.isfinite(rew[k]) and rew[k] < 0 for k in keys)        # -1 (-100 for the ambulance) per step until all passed
    assert rew["a0"] < rew["a1"]
