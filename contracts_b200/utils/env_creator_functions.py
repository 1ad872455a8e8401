"""Drop-in for the reference's utils/env_creator_functions.py: same `env_creator(name, config)` tags,
`get_base_env_tag`, and (when ray is installed) the same `register_env` calls (:12-60)."""
from ..contract import contract_list  # noqa: F401  (the reference module imports it for its side effect)
from ..environments.gridworld import CleanupEnv, HarvestEnv
from ..environments.two_stage_train import (JointEnv, NegotiationSolver, SeparateContractNegotiateStage,
                                            SeparateContractSubgameStage)

_OUT_OF_SCOPE = {
    "ContractWrapperCombined": "SeparateContractCombinedStage is unused by every shipped config",
}


def _feature_env(name):
    def make(**config):
        from ..environments import feature_envs
        return getattr(feature_envs, name)(**config)
    return make


def _selfdrive(**config):
    from ..environments.self_driving_car_accelerate import SelfAcceleratingCarEnv
    return SelfAcceleratingCarEnv(**config)


_CREATORS = {
    "SelfDrive": _selfdrive,
    "Harvest": _feature_env("HarvestFeatures"),
    "HarvestNew": HarvestEnv,
    "Cleanup": _feature_env("CleanupFeatures"),
    "CleanupNew": CleanupEnv,
    "ContractWrapperNegotiate": SeparateContractNegotiateStage,
    "ContractWrapperSubgame": SeparateContractSubgameStage,
    "NegotiationSolver": NegotiationSolver,
    "JointEnv": JointEnv,
}
TAGS = list(_CREATORS) + list(_OUT_OF_SCOPE)


def env_creator(name, config):
    if name in _CREATORS:
        return _CREATORS[name](**config)
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError("%s: %s" % (name, _OUT_OF_SCOPE[name]))
    raise ValueError("Environment not found")


def get_base_env_tag(arg_dict):
    tags = {"selfdrive": "SelfDrive", "harvest": "Harvest", "harvest_new": "HarvestNew", "cleanup": "Cleanup",
            "cleanup_new": "CleanupNew"}
    env = arg_dict.get("environment")
    assert env in tags
    return tags[env]


try:  # pragma: no cover - ray is not part of this image
    from ray.tune.registry import register_env
    for _tag in _CREATORS:
        register_env(_tag, lambda config, _t=_tag: env_creator(_t, config))
except Exception:  # noqa: BLE001
    pass
