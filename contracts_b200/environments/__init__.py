"""Drop-in environment classes: same names, constructor kwargs and RLlib MultiAgentEnv dict API as the
reference's environments/ package, backed by the sm_100a kernels."""
from .gridworld import CleanupEnv, HarvestEnv  # noqa: F401
from .two_stage_train import (SeparateContractEnv, SeparateContractNegotiateStage,  # noqa: F401
                              SeparateContractSubgameStage)
