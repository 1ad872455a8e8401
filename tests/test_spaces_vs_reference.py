"""Container-only (needs /root/reference), CPU: the drop-in classes expose the SAME gym spaces as the reference's classes for
every constructor variant RLlib's policy builder reads them from (`observation_space`, `action_space`, the joint / global /
concatenated / continuous variants; cleanup_new.py:90-169, harvest_new.py:85-130, cleanup_features.py:60-75,
self_driving_car_accelerate.py:36-45, two_stage_train.py:48-60,503-523).  No device is touched: the drop-in objects create
their device batch lazily."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not os.path.isdir("/root/reference/environments"),
                                 reason="reference tree not present (GPU box): container-only test")]

ATTRS = ("observation_space", "action_space", "global_action_space", "global_observation_space",
         "concatenated_observation_space", "continuous_action_space")


def desc(sp):
    t = type(sp).__name__
    if t == "Dict":
        return ("Dict", {k: desc(sp[k]) for k in sorted(sp.keys())})
    if t == "Box":
        return ("Box", tuple(sp.shape), np.asarray(sp.low, dtype=np.float64).tolist(), np.asarray(sp.high, dtype=np.float64).tolist(),
                str(np.dtype(sp.dtype)))
    if t == "Discrete":
        return ("Discrete", int(sp.n))
    if t == "MultiDiscrete":
        return ("MultiDiscrete", np.asarray(sp.nvec).tolist())
    raise AssertionError("unexpected space type %s" % t)


def creators():
    from oracle import ref_stubs
    ref_stubs.install(ref_stubs.reference_root())
    from utils.env_creator_functions import env_creator as ref_creator, get_base_env_tag as ref_tag
    import contract.contract_list as ref_cl
    from contracts_b200.utils.env_creator_functions import env_creator as my_creator, get_base_env_tag as my_tag
    import contracts_b200.contract.contract_list as my_cl
    return ref_creator, ref_cl, ref_tag, my_creator, my_cl, my_tag


BASES = [("CleanupNew", dict(num_agents=3, env_params={}, image_obs=True)),
         ("CleanupNew", dict(num_agents=8, env_params={}, image_obs=True, one_hot_id=True, disable_firing=False)),
         ("CleanupNew", dict(num_agents=2, env_params={}, image_obs=False)),
         ("HarvestNew", dict(num_agents=4, env_params={}, image_obs=True)),
         ("HarvestNew", dict(num_agents=5, env_params={}, image_obs=True, one_hot_id=True)),
         ("HarvestNew", dict(num_agents=4, env_params={}, image_obs=False, disable_firing=False)),
         ("Cleanup", dict(num_agents=5)), ("Harvest", dict(num_agents=3)), ("SelfDrive", dict(num_agents=4)),
         ("SelfDrive", dict(num_agents=2, low_bound=-5.0, high_bound=7.5))]
CONTRACT_OF = {"CleanupNew": "CleanupContract", "HarvestNew": "HarvestFeaturemodLocalContract", "Cleanup": "CleanupContract",
               "Harvest": "HarvestFeaturemodLocalContract", "SelfDrive": "SelfdriveContractDistprop"}


@pytest.mark.parametrize("tag,cfg", BASES, ids=["%s-%d" % (t, i) for i, (t, c) in enumerate(BASES)])
def test_base_env_spaces_match_reference(tag, cfg):
    ref_creator, _, _, my_creator, _, _ = creators()
    r, m = ref_creator(tag, dict(cfg)), my_creator(tag, dict(cfg))
    seen = 0
    for attr in ATTRS:
        if hasattr(r, attr):
            assert hasattr(m, attr), "%s: drop-in class lacks %s" % (tag, attr)
            assert desc(getattr(m, attr)) == desc(getattr(r, attr)), "%s %r: %s differs" % (tag, cfg, attr)
            seen += 1
    assert seen >= 2
    assert m.num_agents == r.num_agents


@pytest.mark.parametrize("tag,cfg", BASES, ids=["%s-%d" % (t, i) for i, (t, c) in enumerate(BASES)])
def test_contract_wrapper_spaces_match_reference(tag, cfg):
    """SeparateContractEnv.__init__ (two_stage_train.py:48-60): the observation gains the `contract` Box (image envs) or two
    more entries (flat observations); the contract's own space comes from contract_list.py."""
    ref_creator, ref_cl, _, my_creator, my_cl, _ = creators()
    n = cfg["num_agents"]
    conv = bool(cfg.get("image_obs"))
    out = []
    for creator, cl in ((ref_creator, ref_cl), (my_creator, my_cl)):
        base = creator(tag, dict(cfg))
        contract = getattr(cl, CONTRACT_OF[tag])(n)
        env = creator("ContractWrapperSubgame", dict(num_agents=n, base_env=base, contract=contract, convolutional=conv))
        out.append((desc(env.observation_space), desc(env.action_space), desc(contract.contract_space)))
    assert out[1] == out[0], "%s %r" % (tag, cfg)


@pytest.mark.parametrize("kwargs", [dict(global_obs=True), dict(concatenated_obs=True)], ids=["global", "concatenated"])
@pytest.mark.parametrize("tag,n", [("CleanupNew", 2), ("HarvestNew", 4)])
def test_joint_env_spaces_match_reference(tag, n, kwargs):
    ref_creator, _, _, my_creator, _, _ = creators()
    out = []
    for creator in (ref_creator, my_creator):
        base = creator(tag, dict(num_agents=n, env_params={}, image_obs=True))
        env = creator("JointEnv", dict(base_env=base, num_agents=n, **kwargs))
        out.append((desc(env.observation_space), desc(env.action_space)))
    assert out[1] == out[0]


def test_base_env_tags_match_reference():
    _, _, ref_tag, _, _, my_tag = creators()
    for name in ("selfdrive", "harvest", "harvest_new", "cleanup", "cleanup_new"):
        assert my_tag({"environment": name}) == ref_tag({"environment": name})
    with pytest.raises(AssertionError):
        ref_tag({"environment": "nope"})
    with pytest.raises(AssertionError):
        my_tag({"environment": "nope"})
